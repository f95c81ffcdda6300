"""Not collected (no test_ prefix) until verified on a GPU: backward modes through the fused-kernel cost functions."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from test_gpu_backward import _golden_module

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("solver", ["dense", "sparse"])
@pytest.mark.parametrize("case", ["plain_implicit_gn", "plain_unroll_lm", "welsch_implicit_lm", "welsch_unroll_lm"])
def test_backward_through_fused_kernel_cost_functions(case, solver):
    """Between / Difference (+ Welsch wrapper) objectives in the backward modes: on the tape these cost functions run through
    their torch restatements (core.CostFunction._torch_error, lie_torch.py), the no-grad iterations through the fused kernels.
    Gradients w.r.t. every edge weight, the prior weight and the robust log-radius against the reference's own autograd
    (torchlie analytic Jacobians): tests/golden/backward_pgo_kat.npz."""
    from helpers import load
    G = _golden_module()
    g = load("backward_pgo_kat")
    skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
        linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization)
    P, grads = G.backward_pgo_run(th, torch, case, device="cuda", solver_kwargs=skw)
    np.testing.assert_allclose(P.cpu().numpy(), g[case + "_poses"], rtol=1e-7, atol=1e-8)
    scale = max(np.abs(g[case + "_grad_" + k]).max() for k in grads)
    for k, v in grads.items():
        ref = g[case + "_grad_" + k]
        assert np.abs(v.cpu().numpy() - ref).max() <= 2e-6 * scale, (k, np.abs(v.cpu().numpy() - ref).max(), scale)


