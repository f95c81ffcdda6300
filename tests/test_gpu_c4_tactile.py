"""BASELINE.json config C4 at its real size on the GPU: the planar-pushing / tactile pose-estimation cost set (SE2 object and effector
poses over T = 25 time steps: QuasiStaticPushingPlanar, EffectorObjectContactPlanar with a bilinear SDF, MovingFrameBetween, SE2 priors;
theseus/embodied/motionmodel/quasi_static_pushing_planar.py, embodied/collision/eff_obj_contact.py,
embodied/measurements/moving_frame_between.py), batch 512, LevenbergMarquardt with adaptive + ellipsoidal damping, and the end-to-end
gradient check in IMPLICIT backward mode (the pattern of tests/theseus_tests/optimizer/nonlinear/test_backwards.py:155-214: gradients of
an outer loss on the solution w.r.t. learnable cost-model parameters).  Golden: tests/golden/tactile_c4_kat.npz, produced by the reference
itself (make_golden.py tactile_c4: dense solver on the CPU).  Tolerances are written in the asserts."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from helpers import load, decisive_iterations
from test_gpu_backward import _golden_module

pytestmark = pytest.mark.gpu
LM = dict(damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)
_KEYS = ("obj", "eff", "eff_meas", "mfb_meas", "c_square", "eff_radius", "sdf", "sdf_origin", "sdf_cell")


def _inputs(g):
    return {k: torch.from_numpy(g[k]) for k in _KEYS}


@pytest.mark.parametrize("solver", ["dense", "sparse_lane", "sparse_front"])
def test_c4_lm_trace_at_batch_512(solver):
    G, g = _golden_module(), load("tactile_c4_kat")
    assert g["obj"].shape[:2] == (25, 512)
    objective, objs, effs, leaves = G.tactile_problem(th, torch, _inputs(g), device="cuda")
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
        linear_solver_kwargs=dict(layout=solver.split("_")[1]))
    iters = g["trace_err"].shape[0]
    opt = th.LevenbergMarquardt(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    errs, deltas = [], []

    def cb(optimizer, info, delta, it):
        errs.append(info.last_err.cpu().numpy().copy()); deltas.append(delta.cpu().numpy().copy())
    with torch.no_grad():
        np.testing.assert_allclose(objective.error_metric().cpu().numpy(), g["err0"], rtol=1e-10)
        opt.optimize(end_iter_callback=cb, **LM)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-7)
    d0 = g["trace_delta"][0]
    rel = np.linalg.norm(deltas[0] - d0, axis=1) / np.linalg.norm(d0, axis=1)
    assert rel.max() < 1e-5, rel.max()          # north star: solution delta within 1e-5 relative of the reference
    np.testing.assert_allclose(np.stack([o.tensor.cpu().numpy() for o in objs], 0), g["final_obj"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("solver", ["dense", "sparse_front"])
def test_c4_implicit_gradients_at_batch_512(solver):
    G, g = _golden_module(), load("tactile_c4_kat")
    objective, objs, effs, leaves = G.tactile_problem(th, torch, _inputs(g), device="cuda")
    for v in leaves.values():
        v.tensor.requires_grad_(True)
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout="front"))
    opt = th.LevenbergMarquardt(objective, max_iterations=8, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    sol, info = th.TheseusLayer(opt).forward({v.name: v.tensor.clone() for v in objs + effs}, optimizer_kwargs=dict(LM, backward_mode="implicit"))
    gen = torch.Generator().manual_seed(5)
    P = torch.stack([sol[o.name] for o in objs], 0)
    (P * torch.randn(P.shape, generator=gen, dtype=torch.float64).cuda()).sum().backward()
    for k, v in leaves.items():
        ref = g["grad_" + k]
        got = v.tensor.grad.cpu().numpy()
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (k, got, ref)
