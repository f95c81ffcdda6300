"""Dense batched Gram kernel (csrc/thb_gram_dense.cu: TMA-staged tiles + FP64 tensor pipe) against torch's fp64 bmm, the operation it
replaces (theseus/optimizer/dense_linearization.py:58-62 `At.bmm(A)`), through the C ABI; and through DenseLinearization for an
AutoDiffCostFunction objective whose Jacobian is genuinely dense.  Tolerance: 1e-12 relative to the largest entry (fp64 products of
O(m) terms summed in a different order)."""
import ctypes as C

import numpy as np
import pytest
import torch

import theseus_b200 as th
from theseus_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,m,n", [(3, 100, 64), (2, 257, 130), (1, 33, 70), (5, 64, 16), (4, 1000, 200), (2, 31, 2)])
def test_gram_dense_matches_bmm(B, m, n):
    g = torch.Generator().manual_seed(B * 1000 + m + n)
    A = torch.randn(B, m, n, generator=g, dtype=torch.float64).cuda()
    out = torch.full((B, n, n), float("nan"), dtype=torch.float64, device="cuda")
    lib = _lib.load()
    _lib.check(lib.thb_gram_dense_f64(_lib.ptr(A), _lib.ptr(out), B, m, n, _lib.stream_ptr()), "gram_dense")
    ref = A.transpose(1, 2) @ A
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max() <= 1e-12 * ref.abs().max()
    assert torch.equal(out, out.transpose(1, 2))          # mirrored tiles: exactly symmetric


def test_gram_dense_refuses_odd_n():
    A = torch.randn(2, 40, 7, dtype=torch.float64).cuda()
    out = torch.empty(2, 7, 7, dtype=torch.float64, device="cuda")
    assert _lib.load().thb_gram_dense_f64(_lib.ptr(A), _lib.ptr(out), 2, 40, 7, _lib.stream_ptr()) == -2   # THB_ERR_UNSUPPORTED


def test_dense_linearization_routes_dense_jacobians_to_the_tma_kernel(monkeypatch):
    """One Vector variable of 24 dof, AutoDiff cost of dim 300 (a random linear-plus-cubic model): J is dense [B, 300, 24]; AtA from the
    TMA kernel must equal the block-Gram route (THB_DENSE_GRAM=0) and J^T J."""
    torch.manual_seed(3)
    B, m, n = 6, 300, 24
    W = torch.randn(m, n, dtype=torch.float64).cuda()

    def err_fn(optim_vars, aux_vars):
        y = torch.einsum("...n,...mn->...m", optim_vars[0].tensor, aux_vars[0].tensor)   # (works batched [B,n] x [1,m,n] and under vmap)
        return y + 0.1 * y ** 3

    outs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("THB_DENSE_GRAM", flag)
        x = th.Vector(tensor=torch.linspace(-1, 1, B * n, dtype=torch.float64).view(B, n).cuda(), name="x")
        Wv = th.Variable(W.unsqueeze(0), name="W")
        objective = th.Objective(dtype=torch.float64)
        cw = th.ScaleCostWeight(torch.tensor(1.0, dtype=torch.float64).cuda())
        objective.add(th.AutoDiffCostFunction([x], err_fn, m, cost_weight=cw, aux_vars=[Wv], name="dense_cost"))
        objective.to("cuda")
        lin = th.DenseLinearization(objective)
        l0 = _lib.total_launches()
        lin.linearize()
        assert lin.engine.dense_jacobian(lin._A_val) == (flag == "1")
        outs[flag] = (lin.AtA.clone(), lin.Atb.clone(), lin.A.clone(), lin.b.clone())
    AtA1, Atb1, A1, b1 = outs["1"]
    AtA0, Atb0, A0, b0 = outs["0"]
    ref = A1.transpose(1, 2) @ A1
    assert (AtA1 - ref).abs().max() <= 1e-12 * ref.abs().max()
    assert (AtA1 - AtA0).abs().max() <= 1e-12 * ref.abs().max()
    assert torch.allclose(Atb1, Atb0, rtol=1e-13, atol=0)
    assert torch.allclose(Atb1.view(B, n), (A1.transpose(1, 2) @ b1.unsqueeze(2)).squeeze(2), rtol=1e-11, atol=1e-11)
