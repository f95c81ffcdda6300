"""The optimizer's host code (engine.py pointer tables, optimizer.py LM / GN loop, solvers) run on the CPU against the reference's
traces, with the CUDA library replaced by its host emulation (tests/simt/: thb_costs.cu, thb_gram.cu, thb_sparse_lane.cu compiled
unchanged, one OS thread per CUDA thread) plus numpy stand-ins for the dense DMMA Cholesky.  Nothing here is a product path: the engine's
CUDA guard and the library loader are replaced explicitly by the test.  What it buys: every change of the host code and of these kernels'
logic is exercised against the reference goldens without a GPU (the GPU suite stays the parity gate)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import theseus_b200 as th
from oracle import nls
from helpers import load, pgo_spec, pgo_objective, lm_kwargs_of, decisive_iterations

HERE = os.path.dirname(os.path.abspath(__file__))


def _emulation_mode():
    spec = importlib.util.spec_from_file_location("emulation_mode", os.path.join(HERE, "simt", "emulation_mode.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def emu_lib():
    return _emulation_mode().load_emulated_lib()


@pytest.fixture
def emulated(monkeypatch, emu_lib):
    _emulation_mode().patch_host(monkeypatch.setattr, emu_lib)
    return emu_lib


def _run(g, solver="dense", **skw_extra):
    method, iters, kw = lm_kwargs_of(g)
    objective, poses = pgo_objective(th, g, device="cpu")
    cls = {"lm": th.LevenbergMarquardt, "gn": th.GaussNewton, "dogleg": th.Dogleg}[method]
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout="lane", **skw_extra))
    opt = cls(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    trace = dict(delta=[], err=[], lam=[])

    def cb(optimizer, info, delta, it):
        trace["delta"].append(delta.numpy().copy())
        trace["err"].append(info.last_err.numpy().copy())
        if method == "lm":
            d = optimizer._damping
            trace["lam"].append(d.numpy().copy() if torch.is_tensor(d) else np.full(delta.shape[0], d))
    layer = th.TheseusLayer(opt)
    inputs = {p.name: p.tensor.clone() for p in poses}
    with torch.no_grad():
        values, info = layer.forward(inputs, optimizer_kwargs=dict(track_err_history=True, end_iter_callback=cb, **kw))
    return method, iters, kw, values, info, trace, poses


@pytest.mark.parametrize("name,solver", [("pgo_small_lm", "dense"), ("pgo_small_gn", "dense"), ("pgo_small_lm_hard", "sparse"),
                                         ("pgo_small_welsch", "dense"), ("pgo_small_geman", "sparse")])
def test_optimizer_traces_on_the_emulated_library(emulated, name, solver):
    g = load(name)
    method, iters, kw, values, info, trace, poses = _run(g, solver)
    ref_err = g["trace_err"]
    mine = np.stack(trace["err"], 0)
    assert mine.shape == ref_err.shape
    np.testing.assert_allclose(mine, ref_err, rtol=1e-8 if method == "lm" else 1e-7)
    spec = pgo_spec(g)
    err0 = nls.error_metric(spec, [v["value"] for v in spec["vars"]])
    np.testing.assert_allclose(info.err_history[:, 0].numpy(), err0, rtol=1e-6)
    k = decisive_iterations(err0, ref_err)
    assert k >= 2
    for it in range(k if method == "lm" else 1):
        dref = g["trace_delta"][it]
        rel = np.linalg.norm(trace["delta"][it] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5, (it, rel)
        if method == "lm":
            np.testing.assert_allclose(trace["lam"][it], g["trace_lam"][it], rtol=1e-12)
    final = np.stack([values[p.name].numpy() for p in poses], 0)
    np.testing.assert_allclose(final, g["poses_final"], rtol=1e-6, atol=1e-6 if k == iters else 5e-4)


def _golden_module():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


LM_TACTILE = dict(damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)


def _tactile_inputs(g):
    return {k: torch.from_numpy(g[k]) for k in ("obj", "eff", "eff_meas", "mfb_meas", "c_square", "eff_radius", "sdf", "sdf_origin", "sdf_cell")}


@pytest.mark.parametrize("solver", ["dense", "sparse"])
def test_tactile_lm_trace_on_the_emulated_library(emulated, solver):
    """Config C4's cost set (QuasiStaticPushingPlanar, EffectorObjectContactPlanar, MovingFrameBetween, SE2 priors: generic route + the fused
    SE2 kernels) -- the CPU twin of tests/test_gpu_zz_first_run.py::test_tactile_lm_trace."""
    G, g = _golden_module(), load("tactile_kat")
    objective, objs, effs, leaves = G.tactile_problem(th, torch, _tactile_inputs(g), device="cpu")
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout="lane"))
    opt = th.LevenbergMarquardt(objective, max_iterations=g["trace_err"].shape[0], step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    errs, deltas = [], []

    def cb(optimizer, info, delta, it):
        errs.append(info.last_err.numpy().copy()); deltas.append(delta.numpy().copy())
    with torch.no_grad():
        np.testing.assert_allclose(objective.error_metric().numpy(), g["err0"], rtol=1e-10)
        opt.optimize(end_iter_callback=cb, **LM_TACTILE)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-7)
    for it in range(decisive_iterations(g["err0"], g["trace_err"])):
        rel = np.linalg.norm(deltas[it] - g["trace_delta"][it], axis=1) / np.linalg.norm(g["trace_delta"][it], axis=1)
        assert rel.max() < 1e-5, (it, rel)
    np.testing.assert_allclose(np.stack([o.tensor.numpy() for o in objs], 0), g["final_obj"], rtol=1e-5, atol=1e-6)


def test_tactile_implicit_gradients_on_the_emulated_library(emulated):
    G, g = _golden_module(), load("tactile_kat")
    objective, objs, effs, leaves = G.tactile_problem(th, torch, _tactile_inputs(g), device="cpu")
    for v in leaves.values():
        v.tensor.requires_grad_(True)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=8, step_size=1.0, abs_err_tolerance=0,
                                rel_err_tolerance=0)
    sol, info = th.TheseusLayer(opt).forward({v.name: v.tensor.clone() for v in objs + effs}, optimizer_kwargs=dict(LM_TACTILE, backward_mode="implicit"))
    gen = torch.Generator().manual_seed(5)
    P = torch.stack([sol[o.name] for o in objs], 0)
    (P * torch.randn(P.shape, generator=gen, dtype=torch.float64)).sum().backward()
    for k, v in leaves.items():
        ref = g["grad_" + k]
        assert np.abs(v.tensor.grad.numpy() - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max()), k


@pytest.mark.parametrize("solver", ["dense", "sparse"])
def test_so2_lm_trace_on_the_emulated_library(emulated, solver):
    """SO2 variables: the THB_VAR_SO2 branch of the retract kernel + Between / Difference on the generic route (CPU twin of the pending GPU test)."""
    G, g = _golden_module(), load("so2_kat")
    objective, vs = G.so2_problem(th, torch, torch.from_numpy(g["lm_thetas0"]), torch.from_numpy(g["lm_meas"]),
                                  [tuple(int(x) for x in e) for e in g["lm_edges"]], g["lm_w_edge"], float(g["lm_w_prior"]))
    iters = g["lm_trace_err"].shape[0]
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout="lane"))
    opt = th.LevenbergMarquardt(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    errs, deltas = [], []

    def cb(optimizer, info, delta, it):
        errs.append(info.last_err.numpy().copy()); deltas.append(delta.numpy().copy())
    with torch.no_grad():
        np.testing.assert_allclose(objective.error_metric().numpy(), g["lm_err0"], rtol=1e-12)
        opt.optimize(end_iter_callback=cb, damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)
    np.testing.assert_allclose(np.stack(errs, 0), g["lm_trace_err"], rtol=1e-9)
    for it in range(iters):
        np.testing.assert_allclose(deltas[it], g["lm_trace_delta"][it], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(np.stack([v.tensor.numpy() for v in vs], 0), g["lm_final"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("solver", ["dense", "sparse"])
def test_custom_variable_ordering_on_the_emulated_library(emulated, solver):
    g = load("pgo_small_lm")
    method, iters, kw = lm_kwargs_of(g)
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout="lane"))
    runs = {}
    for mode in ("default", "reversed"):
        objective, poses = pgo_objective(th, g, device="cpu")
        lkw = {}
        if mode == "reversed":
            order = th.VariableOrdering(objective, default_order=False)
            order.extend(list(reversed(poses)))
            lkw = dict(linearization_kwargs=dict(ordering=order))
        opt = th.LevenbergMarquardt(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw, **lkw)
        errs, deltas = [], []

        def cb(optimizer, info, delta, it):
            errs.append(info.last_err.numpy().copy()); deltas.append(delta.numpy().copy())
        with torch.no_grad():
            opt.optimize(end_iter_callback=cb, **kw)
        if mode == "reversed":
            assert objective.engine().custom_ordering == tuple(p.name for p in reversed(poses))
        runs[mode] = (np.stack(errs, 0), np.stack(deltas, 0), np.stack([p.tensor.numpy() for p in poses], 0))
    np.testing.assert_allclose(runs["reversed"][0], g["trace_err"], rtol=1e-8)
    N = len(runs["default"][2])
    d_def = runs["default"][1].reshape(iters, -1, N, 6)
    d_rev = runs["reversed"][1].reshape(iters, -1, N, 6)[:, :, ::-1]
    rel = np.linalg.norm((d_def - d_rev).reshape(iters, -1), axis=1) / np.linalg.norm(d_def.reshape(iters, -1), axis=1)
    assert rel[:2].max() < 1e-6, rel
    np.testing.assert_allclose(runs["reversed"][2], g["poses_final"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("name", ["pgo_small_welsch", "pgo_small_geman"])
def test_batched_torch_route_inside_the_engine(emulated, name, monkeypatch):
    g = load(name)
    objective, poses = pgo_objective(th, g, device="cpu")
    eng = objective.engine()
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("THB_BATCHED_TORCH_ROUTE", flag)
        A, b = eng.linearize_sparse_differentiable()
        A2, b2 = eng.linearize_sparse()
        outs[flag] = (A.detach().numpy(), b.detach().numpy(), A2.numpy().copy(), b2.numpy().copy(), eng.error_metric().numpy().copy())
    for x0, x1 in zip(outs["0"], outs["1"]):
        np.testing.assert_allclose(x1, x0, rtol=1e-12, atol=1e-13 * np.abs(x0).max())
    np.testing.assert_allclose(outs["1"][0], outs["1"][2], rtol=1e-7, atol=1e-9 * np.abs(outs["1"][2]).max())   # taped (autodiff) == fused-kernel (analytic) values


def test_returned_solutions_are_not_overwritten_by_the_next_forward(emulated):
    """TheseusLayer.forward hands out a snapshot of the engine-owned variable pool: a second forward (new inputs, same objective) must
    leave the tensors returned by the first one untouched -- the reference rebinds fresh tensors (core/variable.py:42-72)."""
    g = load("pgo_small_lm")
    method, iters, kw = lm_kwargs_of(g)
    objective, poses = pgo_objective(th, g, device="cpu")
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=3, step_size=1.0)
    layer = th.TheseusLayer(opt)
    inputs = {p.name: p.tensor.clone() for p in poses}
    with torch.no_grad():
        sol1, _ = layer.forward(inputs, optimizer_kwargs=kw)
        keep = {k: v.clone() for k, v in sol1.items()}
        inputs2 = {k: v.clone() for k, v in inputs.items()}
        inputs2[poses[1].name] = sol1[poses[1].name].clone()
        sol2, _ = layer.forward(inputs2, optimizer_kwargs=kw)
    for k in keep:
        assert torch.equal(sol1[k], keep[k]), k
        assert sol1[k].data_ptr() != sol2[k].data_ptr()
    assert any(not torch.equal(sol1[k], sol2[k]) for k in keep)


def test_error_vector_of_robust_costs_matches_the_error_metric(emulated):
    """objective.py:562-641: error_metric() == 0.5 * ||error()||^2, also with RobustCostFunctions (whose weighted error is
    sqrt(rho / dim + eps) per entry, not the sqrt(rho')-rescaled residual of the linearization)."""
    g = load("pgo_small_welsch")
    objective, poses = pgo_objective(th, g, device="cpu")
    with torch.no_grad():
        e = objective.error()
        m = objective.error_metric()
    assert e.shape[0] == m.shape[0] and e.shape[1] == objective.dim()
    np.testing.assert_allclose(0.5 * (e ** 2).sum(dim=1).numpy(), m.numpy(), rtol=1e-10)
