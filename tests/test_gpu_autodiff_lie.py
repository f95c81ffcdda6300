"""AutoDiffCostFunction over Lie-group variables: torch.func Euclidean Jacobians projected onto the tangent space
(cost_function.py:343-393, v.project(jac, is_sparse=True)), scattered into the batched-CSR Jacobian, then the fused solve / retract
kernels -- against the reference's own A_val / b and LM trace (tests/golden/autodiff_lie.npz)."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from helpers import load, decisive_iterations
from test_gpu_backward import _golden_module

pytestmark = pytest.mark.gpu


def _objective(g):
    G = _golden_module()
    inputs = {k: torch.from_numpy(g[k]) for k in ("T3", "T2", "p", "q", "p2", "q2")}
    objective = G.autodiff_lie_problem(th, torch, inputs)
    objective.to("cuda")
    return objective


def test_projected_jacobians_match_reference():
    g = load("autodiff_lie")
    objective = _objective(g)
    lin = th.SparseLinearization(objective)
    objective.update()
    lin.linearize()
    np.testing.assert_allclose(lin.A_val.cpu().numpy(), g["A_val0"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(lin.b.cpu().numpy(), g["b0"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("solver", ["dense", "sparse"])
def test_lm_trace_matches_reference(solver):
    g = load("autodiff_lie")
    objective = _objective(g)
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization)
    opt = th.LevenbergMarquardt(objective, max_iterations=g["trace_err"].shape[0], step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, **skw)
    deltas, errs = [], []

    def cb(optimizer, info, delta, it):
        deltas.append(delta.cpu().numpy().copy()); errs.append(info.last_err.cpu().numpy().copy())

    with torch.no_grad():
        err0 = objective.error_metric().cpu().numpy()
        opt.optimize(end_iter_callback=cb, damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-8)
    k = decisive_iterations(err0, g["trace_err"])
    assert k >= 2
    for it in range(k):
        dref = g["trace_delta"][it]
        rel = np.linalg.norm(deltas[it] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5, (it, rel)
    np.testing.assert_allclose(objective.get_optim_var("T3").tensor.cpu().numpy(), g["final_T3"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(objective.get_optim_var("T2").tensor.cpu().numpy(), g["final_T2"], rtol=1e-6, atol=1e-6)
