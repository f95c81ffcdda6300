"""SO3 cost kernels + retract, the batched-CSR helpers (extlib/mat_mult.cu equivalents, checked against scipy like
tests/theseus_tests/extlib/test_mat_mult.py:14-146, atol 1e-10) and the optimizer info bookkeeping."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import theseus_b200 as th
from theseus_b200 import _lib
from oracle import nls, lie
from helpers import load, pgo_objective, lm_kwargs_of

pytestmark = pytest.mark.gpu


def test_so3_costs_and_retract_vs_reference():
    g = load("so3_kat")
    dt = torch.float64
    X0 = th.SO3(tensor=torch.from_numpy(g["X0"]), name="X0")
    X1 = th.SO3(tensor=torch.from_numpy(g["X1"]), name="X1")
    Z = th.SO3(tensor=torch.from_numpy(g["Z"]), name="Z")
    obj = th.Objective(dtype=dt)
    obj.add(th.Between(X0, X1, Z, th.DiagonalCostWeight(th.Variable(torch.from_numpy(g["w"]), name="w")), name="b"))
    obj.add(th.Difference(X0, Z, th.ScaleCostWeight(torch.tensor(1.3, dtype=dt)), name="l"))
    obj.to("cuda")
    lin = th.DenseLinearization(obj)
    lin.linearize()
    A = lin.A.cpu().numpy()
    np.testing.assert_allclose(A[:, 0:3, 0:3], g["between_J0"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(A[:, 0:3, 3:6], g["between_J1"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(A[:, 3:6, 0:3], g["local_J"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(-lin.b.cpu().numpy(), np.concatenate([g["between_e"], g["local_e"]], 1), rtol=1e-10, atol=1e-12)
    # oracle agrees with the reference too
    jo, eo = nls.between_error_jacobians("SO3", g["X0"], g["X1"], g["Z"])
    jo, eo = nls.weight_jacobians_error(("diag", g["w"]), jo, eo)
    np.testing.assert_allclose(eo, g["between_e"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lie.so3_retract(g["X0"], g["delta"]), g["retract"], rtol=1e-12, atol=1e-14)
    # retract kernel: delta for [X0 | X1]
    eng = obj.engine()
    eng.adopt_optim_vars()
    tmp = [X0.copy(new_name="X0"), X1.copy(new_name="X1")]
    delta = torch.zeros(g["delta"].shape[0], 6, dtype=dt, device="cuda")
    delta[:, :3] = torch.from_numpy(g["delta"]).cuda()
    obj.retract_vars_sequence(delta, tmp)
    np.testing.assert_allclose(tmp[0].tensor.cpu().numpy(), g["retract"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(tmp[1].tensor.cpu().numpy(), g["X1"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("B,rows,cols,fill", [(1, 50, 30, 0.1), (32, 50, 30, 0.1), (8, 200, 70, 0.05)])
def test_mat_vec_and_tmat_vec_vs_scipy(B, rows, cols, fill):
    rng = np.random.default_rng(rows + B)
    M = sp.random(rows, cols, density=fill, format="csr", random_state=1, dtype=np.float64)
    M.sort_indices()
    lib = _lib.load()
    rp = torch.from_numpy(M.indptr.astype(np.int64)).cuda()
    ci = torch.from_numpy(M.indices.astype(np.int64)).cuda()
    vals = rng.standard_normal((B, M.nnz))
    A_val = torch.from_numpy(vals).cuda()
    v = rng.standard_normal((B, cols)); u = rng.standard_normal((B, rows))
    y = torch.empty(B, rows, dtype=torch.float64, device="cuda")
    yt = torch.empty(B, cols, dtype=torch.float64, device="cuda")
    _lib.check(lib.thb_mat_vec_f64(B, rows, cols, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(A_val), _lib.ptr(torch.from_numpy(v).cuda()), _lib.ptr(y), _lib.stream_ptr()), "mat_vec")
    _lib.check(lib.thb_tmat_vec_f64(B, rows, cols, _lib.ptr(rp), _lib.ptr(ci), _lib.ptr(A_val), _lib.ptr(torch.from_numpy(u).cuda()), _lib.ptr(yt), _lib.stream_ptr()), "tmat_vec")
    for b in range(B):
        Mb = sp.csr_matrix((vals[b], M.indices, M.indptr), shape=(rows, cols))
        np.testing.assert_allclose(y[b].cpu().numpy(), Mb @ v[b], atol=1e-10)
        np.testing.assert_allclose(yt[b].cpu().numpy(), Mb.T @ u[b], atol=1e-10)


def test_info_bookkeeping_best_solution_and_histories():
    g = load("pgo_small_lm_hard")
    method, iters, kw = lm_kwargs_of(g)
    objective, poses = pgo_objective(th, g)
    opt = th.LevenbergMarquardt(objective, max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0)
    with torch.no_grad():
        info = opt.optimize(track_best_solution=True, track_err_history=True, track_state_history=True, **kw)
    B = g["poses0"].shape[1]
    assert info.err_history.shape == (B, iters + 1) and torch.isfinite(info.err_history).all()
    np.testing.assert_allclose(info.err_history[:, 1:].numpy().T, g["trace_err"], rtol=1e-6)   # histories are kept in fp32 on the CPU like the reference
    assert set(info.best_solution) == {p.name for p in poses}
    name = poses[3].name
    assert info.state_history[name].shape == (B, 3, 4, iters + 1)
    np.testing.assert_allclose(info.state_history[name][..., 0].numpy(), g["poses0"][3], rtol=1e-6)
    np.testing.assert_allclose(info.best_err.cpu().numpy(), g["trace_err"].min(axis=0), rtol=1e-8)
    assert (info.converged_iter == -1).all() and all(s == th.NonlinearOptimizerStatus.MAX_ITERATIONS for s in info.status)


@pytest.mark.parametrize("solver", ["dense", "sparse_lane"])
@pytest.mark.parametrize("name", ["pgo_small_lm", "pgo_small_gn", "pgo64_lm"])
def test_cuda_graph_mode_is_bitwise_identical_to_eager(name, solver):
    """cuda_graph=True replays the captured iteration body; same kernels, same order -> identical bits; also across repeated
    optimize() calls (graph reuse) and with the all-rejected / non-PD bookkeeping read after each replay."""
    from helpers import load, pgo_objective, lm_kwargs_of
    g = load(name)
    method, iters, kw = lm_kwargs_of(g)
    cls = th.LevenbergMarquardt if method == "lm" else th.GaussNewton
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout="lane"))
    res = {}
    for mode in (False, True):
        objective, poses = pgo_objective(th, g)
        opt = cls(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, cuda_graph=mode, **skw)
        layer = th.TheseusLayer(opt)
        inputs = {p.name: p.tensor.clone() for p in poses}
        outs = []
        for rep in range(2):   # second call re-uses the captured graph
            with torch.no_grad():
                vals, info = layer.forward(inputs, optimizer_kwargs=dict(kw, track_err_history=True))
            outs.append((torch.stack([vals[p.name] for p in poses]).cpu().numpy().copy(), info.err_history.numpy().copy(),
                         info.last_err.cpu().numpy().copy()))
        res[mode] = outs
        if mode:
            assert opt._graph is not None
    for rep in range(2):
        for a, b in zip(res[False][rep], res[True][rep]):
            assert np.array_equal(a, b)
    assert np.array_equal(res[True][0][0], res[True][1][0])
