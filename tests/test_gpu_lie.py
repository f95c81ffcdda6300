"""CUDA Lie kernels (through the C ABI) vs the reference's own outputs (golden) and vs the oracle on random inputs."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from oracle import lie
from helpers import load

pytestmark = pytest.mark.gpu


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_se3_kernels_vs_reference_golden(dt):
    g = load("lie_kat")
    p = f"se3_{dt}_"
    tol = dict(rtol=1e-9, atol=1e-10) if dt == "f64" else dict(rtol=2e-3, atol=2e-4)
    xi = _t(g[p + "tangent"])
    G = th.SE3.exp_map(xi)
    np.testing.assert_allclose(G.tensor.cpu().numpy(), g[p + "exp"], **(dict(rtol=1e-10, atol=1e-12) if dt == "f64" else dict(rtol=2e-4, atol=2e-5)))
    Gref = th.SE3(tensor=_t(g[p + "exp"]))
    jl = []
    lx = Gref.log_map(jacobians=jl)
    np.testing.assert_allclose(lx.cpu().numpy(), g[p + "log"], **tol)
    if dt == "f64":
        np.testing.assert_allclose(jl[0].cpu().numpy(), g[p + "jlog"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(Gref.adjoint().cpu().numpy(), g[p + "adj"], **tol)
    np.testing.assert_allclose(Gref.inverse().tensor.cpu().numpy(), g[p + "inv"], **tol)
    other = th.SE3(tensor=_t(g[p + "other"]))
    np.testing.assert_allclose(Gref.compose(other).tensor.cpu().numpy(), g[p + "compose"], **tol)


def test_se3_kernels_vs_oracle_random_large():
    rng = np.random.default_rng(11)
    N = 100_000
    xi = rng.standard_normal((N, 6))
    xi[: N // 4, 3:] *= 1e-3  # exercise the near-zero branches
    G = th.SE3.exp_map(_t(xi))
    np.testing.assert_allclose(G.tensor.cpu().numpy(), lie.se3_exp(xi), rtol=1e-10, atol=1e-12)
    jl = []
    lx = G.log_map(jacobians=jl)
    Jo, lo = lie.se3_jlog(G.tensor.cpu().numpy())
    # log is ill-conditioned as the rotation angle approaches pi (d log ~ 1/(pi - theta)): identical formulas fed
    # with R differing by 1 ulp differ by ~1e-16/(pi-theta)^2; keep the tight bar away from pi, a loose one near it
    ang = np.linalg.norm(lo[:, 3:], axis=1)
    far = ang < np.pi - 0.05
    np.testing.assert_allclose(lx.cpu().numpy()[far], lo[far], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(jl[0].cpu().numpy()[far], Jo[far], rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(lx.cpu().numpy()[~far], lo[~far], rtol=1e-4, atol=1e-4)
    # size-independent property: log(exp(x)) == x away from pi
    m = np.linalg.norm(xi[:, 3:], axis=1) < 3.0
    # NOT to 1e-9: below near_zero (theta < 5e-3) the reference's exp uses sin(t)/t ~ (1+cos t)/2 (so3_impl.py:231-233),
    # a theta^2/12 ~ 2e-6 relative approximation of the translation part, which this implementation reproduces
    np.testing.assert_allclose(lx.cpu().numpy()[m], xi[m], rtol=1e-5, atol=5e-6)
    big = m & (np.linalg.norm(xi[:, 3:], axis=1) > 1e-2)
    np.testing.assert_allclose(lx.cpu().numpy()[big], xi[big], rtol=1e-8, atol=1e-9)


def test_cpu_tensor_fails_loudly():
    with pytest.raises(RuntimeError, match="CUDA"):
        th.SE3.exp_map(torch.zeros(2, 6, dtype=torch.float64))


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_so3_kernels_vs_reference_golden(dt):
    """Stand-alone SO3 kernels (thb_so3_*) against torchlie.functional outputs incl. the reference's angle sweep (lie_kat.npz)."""
    g = load("lie_kat")
    tdt = torch.float64 if dt == "f64" else torch.float32
    tol = dict(rtol=1e-9, atol=1e-10) if dt == "f64" else dict(rtol=2e-3, atol=2e-4)
    P = lambda k: torch.from_numpy(g[f"so3_{dt}_{k}"]).to(tdt).cuda()
    X = th.SO3.exp_map(P("tangent"))
    np.testing.assert_allclose(X.tensor.cpu().numpy(), g[f"so3_{dt}_exp"], **(dict(rtol=1e-10, atol=1e-12) if dt == "f64" else dict(rtol=2e-4, atol=2e-5)))
    Y = th.SO3(tensor=P("exp"))
    jac = []
    w = Y.log_map(jacobians=jac)
    np.testing.assert_allclose(w.cpu().numpy(), g[f"so3_{dt}_log"], **tol)
    if dt == "f64":
        np.testing.assert_allclose(jac[0].cpu().numpy(), g[f"so3_{dt}_jlog"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(Y.adjoint().cpu().numpy(), g[f"so3_{dt}_adj"], **tol)
    np.testing.assert_allclose(Y.inverse().tensor.cpu().numpy(), g[f"so3_{dt}_inv"], **tol)
    np.testing.assert_allclose(Y.compose(th.SO3(tensor=P("other"))).tensor.cpu().numpy(), g[f"so3_{dt}_compose"], **tol)


def test_se2_kernels_vs_reference_golden():
    """Stand-alone SE2 kernels (thb_se2_*) against theseus.geometry.SE2 outputs (se2_kat.npz)."""
    g = load("se2_kat")
    P = lambda k: torch.from_numpy(g[k]).cuda()
    np.testing.assert_allclose(th.SE2.exp_map(P("tangent")).tensor.cpu().numpy(), g["exp"], rtol=1e-12, atol=1e-14)
    Y = th.SE2(tensor=P("exp"))
    jac = []
    np.testing.assert_allclose(Y.log_map(jacobians=jac).cpu().numpy(), g["log"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(jac[0].cpu().numpy(), g["jlog"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(Y.adjoint().cpu().numpy(), g["adj"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(Y.inverse().tensor.cpu().numpy(), g["inv"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(Y.compose(th.SE2(tensor=P("other"))).tensor.cpu().numpy(), g["compose"], rtol=1e-12, atol=1e-14)
