"""SE2 (C4's group) on the GPU: Between / Difference linearization and the LM trace of a 2-D pose graph vs the reference."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from helpers import load, se2_pg_objective

pytestmark = pytest.mark.gpu


def test_se2_cost_functions_vs_reference():
    g = load("se2_kat")
    dt = torch.float64
    G = th.SE2(tensor=torch.from_numpy(g["exp"]), name="G")
    H = th.SE2(tensor=torch.from_numpy(g["other"]), name="H")
    Z = th.SE2(tensor=torch.from_numpy(g["Z"]), name="Z")
    obj = th.Objective(dtype=dt)
    obj.add(th.Between(G, H, Z, th.DiagonalCostWeight(th.Variable(torch.from_numpy(g["w"]), name="w")), name="b"))
    obj.add(th.Difference(G, Z, th.ScaleCostWeight(torch.tensor(0.7, dtype=dt)), name="l"))
    obj.to("cuda")
    lin = th.DenseLinearization(obj)
    lin.linearize()
    A = lin.A.cpu().numpy()
    np.testing.assert_allclose(A[:, 0:3, 0:3], g["between_J0"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(A[:, 0:3, 3:6], g["between_J1"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(A[:, 3:6, 0:3], g["local_J"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(-lin.b.cpu().numpy()[:, 0:3], g["between_e"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(-lin.b.cpu().numpy()[:, 3:6], g["local_e"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("solver", ["dense", "sparse"])
def test_se2_pose_graph_lm_trace(solver):
    g = load("se2_kat")
    objective, poses = se2_pg_objective(th, g)
    kw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization)
    opt = th.LevenbergMarquardt(objective, max_iterations=6, abs_err_tolerance=0, rel_err_tolerance=0, **kw)
    deltas, errs = [], []

    def cb(optimizer, info, delta, it):
        deltas.append(delta.cpu().numpy().copy()); errs.append(info.last_err.cpu().numpy().copy())
    with torch.no_grad():
        opt.optimize(end_iter_callback=cb, damping=1e-2, adaptive_damping=True, ellipsoidal_damping=True)
    np.testing.assert_allclose(np.stack(errs, 0), g["pg_trace_err"], rtol=1e-8)
    for it in range(3):
        dref = g["pg_trace_delta"][it]
        rel = np.linalg.norm(deltas[it] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5
    np.testing.assert_allclose(np.stack([p.tensor.cpu().numpy() for p in poses], 0), g["pg_final"], rtol=1e-6, atol=1e-7)
