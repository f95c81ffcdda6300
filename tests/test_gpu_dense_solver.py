"""Batched dense Cholesky factor+solve with fused damping (thb_potrf_potrs_f64) vs the reference's
torch.linalg.cholesky/cholesky_solve outputs (golden) and in the style of
tests/theseus_tests/optimizer/linear/test_dense_solver.py:12-77 (random SPD systems, both damping modes)."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from oracle import nls
from helpers import load

pytestmark = pytest.mark.gpu


def _solver():
    obj = th.Objective(dtype=torch.float64)
    v = th.Vector(tensor=torch.zeros(1, 1, dtype=torch.float64), name="v")
    obj.add(th.Difference(v, th.Vector(tensor=torch.zeros(1, 1, dtype=torch.float64), name="t"), th.ScaleCostWeight(torch.tensor(1.0, dtype=torch.float64))))
    return th.CholeskyDenseSolver(obj)


def test_golden_kat():
    g = load("dense_solver_kat")
    s = _solver()
    for n in (1, 6, 10, 48, 130):
        AtA = torch.from_numpy(g[f"n{n}_AtA"]).cuda()
        Atb = torch.from_numpy(g[f"n{n}_Atb"]).cuda()
        lam = torch.from_numpy(g[f"n{n}_lam"]).cuda()
        for ell, tag in ((True, "ell"), (False, "sph")):
            x = s._apply_damping_and_solve(Atb, AtA, damping=lam, ellipsoidal_damping=ell, damping_eps=1e-8)
            np.testing.assert_allclose(x.cpu().numpy(), g[f"n{n}_x_{tag}"], rtol=1e-8, atol=1e-10)
        x0 = s._apply_damping_and_solve(Atb, AtA)  # no damping
        r = torch.bmm(AtA, x0.unsqueeze(2)) - Atb
        assert r.abs().max() < 1e-8


@pytest.mark.parametrize("B,n", [(32, 10), (7, 127), (5, 128), (5, 129), (3, 300), (2, 700), (2, 1536)])
def test_random_spd_residual(B, n):
    gen = torch.Generator(device="cuda").manual_seed(n)
    A = torch.randn(B, n + 5, n, dtype=torch.float64, device="cuda", generator=gen)
    AtA = A.transpose(1, 2) @ A + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
    xt = torch.randn(B, n, 1, dtype=torch.float64, device="cuda", generator=gen)
    lam = torch.rand(B, dtype=torch.float64, device="cuda", generator=gen) + 0.01
    s = _solver()
    for ell in (True, False):
        D = AtA.clone()
        idx = torch.arange(n, device="cuda")
        if ell:
            D[:, idx, idx] += lam.view(-1, 1) * AtA[:, idx, idx] + 1e-8
        else:
            D[:, idx, idx] += lam.view(-1, 1)
        rhs = D @ xt
        x = s._apply_damping_and_solve(rhs, AtA, damping=lam, ellipsoidal_damping=ell, damping_eps=1e-8)
        rel = (x - xt.squeeze(2)).norm(dim=1) / xt.squeeze(2).norm(dim=1)
        assert rel.max() < 1e-9, rel  # same bar as test_dense_solver.py (1e-4) with fp64 headroom
    # AtA must be untouched (LM reads its diagonal afterwards)
    A2 = A.transpose(1, 2) @ A + 0.1 * torch.eye(n, dtype=torch.float64, device="cuda")
    assert torch.equal(AtA, A2)


def test_scalar_damping_and_oracle():
    rng = np.random.default_rng(2)
    B, n = 4, 60
    M = rng.standard_normal((B, n + 2, n))
    AtA = np.einsum("bij,bik->bjk", M, M) + 0.1 * np.eye(n)
    Atb = rng.standard_normal((B, n, 1))
    s = _solver()
    x = s._apply_damping_and_solve(torch.from_numpy(Atb).cuda(), torch.from_numpy(AtA).cuda(), damping=0.05,
                                   ellipsoidal_damping=False)
    xo = nls.dense_solve(AtA, Atb, damping=0.05, ellipsoidal=False)
    # (n+2) x n Gaussian A: cond(AtA + 0.05 I) ~ 1e4..1e5, so two backward-stable solvers agree to ~cond*eps
    rel = np.linalg.norm(x.cpu().numpy() - xo, axis=1) / np.linalg.norm(xo, axis=1)
    assert rel.max() < 1e-7, rel
    D = AtA + 0.05 * np.eye(n)
    res = np.einsum("bij,bj->bi", D, x.cpu().numpy()) - Atb[:, :, 0]
    # the reference's own criterion for its solvers (extlib/test_baspacho.py:101-114), scaled by ||D|| ||x|| ~ 1e3
    assert np.abs(res).max() < 1e-10 * np.abs(D).sum(axis=1).max() * max(1.0, np.abs(xo).max())


def test_not_positive_definite_raises_runtime_error():
    s = _solver()
    AtA = torch.eye(5, dtype=torch.float64, device="cuda").repeat(3, 1, 1)
    AtA[1, 2, 2] = -1.0
    with pytest.raises(RuntimeError, match="positive-definite"):
        s._apply_damping_and_solve(torch.ones(3, 5, 1, dtype=torch.float64, device="cuda"), AtA)
