"""Bundle adjustment (C3 shape: Reprojection + Difference priors on SE3 cameras / Point3 points) through the product on the
GPU vs the reference's own numbers: CSR structure bit-exact, A_val/b, and the LM trace with the dense and the block-sparse solver."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from oracle import nls
from helpers import load, ba_objective, ba_spec, lm_kwargs_of, decisive_iterations

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["ba_small_lm", "ba_small_huber"])
def test_ba_linearization_matches_reference(name):
    g = load(name)
    objective, cams, pts = ba_objective(th, g)
    lin = th.SparseLinearization(objective)
    assert [v.name for v in lin.ordering] == [str(x) for x in g["order"]]
    assert np.array_equal(lin.A_row_ptr, g["A_row_ptr"]) and np.array_equal(lin.A_col_ind, g["A_col_ind"])
    lin.linearize()
    np.testing.assert_allclose(lin.A_val.cpu().numpy(), g["A_val0"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(lin.b.cpu().numpy(), g["b0"], rtol=1e-9, atol=1e-9)
    spec = ba_spec(g)
    np.testing.assert_allclose(objective.error_metric().cpu().numpy(), nls.error_metric(spec, [v["value"] for v in spec["vars"]]), rtol=1e-12)


@pytest.mark.parametrize("name", ["ba_small_lm", "ba_small_huber", "ba_c3_huber"])   # ba_c3_huber: config C3 at full size (50 x 1000 x 8)
@pytest.mark.parametrize("solver", ["dense", "sparse", "sparse_lane"])
def test_ba_lm_trace(solver, name):
    g = load(name)
    method, iters, kw = lm_kwargs_of(g)
    objective, cams, pts = ba_objective(th, g)
    if solver == "dense":
        opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0)
    else:
        opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                                    max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0,
                                    linear_solver_kwargs=dict(layout="lane" if solver == "sparse_lane" else "item"))
        # minimum degree must eliminate the (many, cheap) points before the cameras
        plan = opt.linear_solver._plan
        first_cam = min(int(plan.pos[i]) for i, v in enumerate(opt.linear_solver.linearization.ordering) if v.dof() == 6)
        assert first_cam >= 30
    deltas, errs, lams = [], [], []

    def cb(optimizer, info, delta, it):
        deltas.append(delta.cpu().numpy().copy()); errs.append(info.last_err.cpu().numpy().copy()); lams.append(optimizer._damping.cpu().numpy().copy())

    with torch.no_grad():
        info = opt.optimize(end_iter_callback=cb, **kw)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-8)
    spec = ba_spec(g)
    err0 = nls.error_metric(spec, [v["value"] for v in spec["vars"]])
    k = decisive_iterations(err0, g["trace_err"])
    for it in range(k):
        dref = g["trace_delta"][it]
        rel = np.linalg.norm(deltas[it] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5, (it, rel)
        np.testing.assert_allclose(lams[it], g["trace_lam"][it], rtol=1e-12)
    final = np.concatenate([v.tensor.cpu().numpy().reshape(v.tensor.shape[0], -1) for v in opt.linear_solver.linearization.ordering], 1)
    np.testing.assert_allclose(final, g["final"], rtol=1e-5, atol=1e-5)
