"""SO2 (theseus/geometry/so2.py): the oracle's restatement and the package's class / torch route against values computed
by the reference (tests/golden/make_golden.py:make_so2 -> so2_kat.npz).  Inside the optimizer SO2 variables are retracted by the fused
retract kernel and their cost functions take the engine's generic route; that half runs on the GPU (tests/test_gpu_zz_first_run.py)."""
import numpy as np
import torch

import theseus_b200 as th
from helpers import load
from oracle import lie, nls


def test_oracle_so2_known_answers():
    g = load("so2_kat")
    X, Y, Z = g["X"], g["Y"], g["Z"]
    np.testing.assert_allclose(lie.so2_exp(g["theta"]), X, rtol=0, atol=1e-15)
    np.testing.assert_allclose(lie.so2_log(X), g["log_X"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(lie.so2_compose(X, Y), g["compose"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(lie.so2_inverse(X), g["inverse"], rtol=0, atol=0)
    np.testing.assert_allclose(lie.so2_adjoint(X), g["adjoint"])
    np.testing.assert_allclose(lie.so2_log(lie.so2_compose(lie.so2_inverse(X), Y)), g["local"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(lie.so2_retract(X, g["retract_delta"]), g["retract"], rtol=0, atol=1e-15)
    jacs, e = nls.between_error_jacobians("SO2", X, Y, Z)
    jacs, e = nls.weight_jacobians_error(("scale", np.full((1, 1), 0.7)), jacs, e)
    np.testing.assert_allclose(e, g["between_e"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(jacs[0], g["between_J0"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(jacs[1], g["between_J1"], rtol=0, atol=1e-14)
    jl, el = nls.local_error_jacobians("SO2", X, Z)
    jl, el = nls.weight_jacobians_error(("scale", np.full((1, 1), 1.3)), jl, el)
    np.testing.assert_allclose(el, g["local_e"], rtol=0, atol=1e-14)
    np.testing.assert_allclose(jl[0], g["local_J"], rtol=0, atol=1e-14)


def _so2_spec(g):
    th0, meas, edges = g["lm_thetas0"], g["lm_meas"], g["lm_edges"]
    spec = dict(dtype=np.dtype(np.float64), vars=[], costs=[])
    for i in range(th0.shape[0]):
        spec["vars"].append(dict(kind="SO2", dof=1, value=lie.so2_exp(th0[i])))
    for e in range(edges.shape[0]):
        spec["costs"].append(dict(kind="between", group="SO2", vars=(int(edges[e, 0]), int(edges[e, 1])), aux=lie.so2_exp(meas[e]),
                                  weight=("scale", np.full((1, 1), float(g["lm_w_edge"][e])))))
    spec["costs"].append(dict(kind="local", group="SO2", vars=(0,), aux=lie.so2_exp(th0[0]), weight=("scale", np.full((1, 1), float(g["lm_w_prior"])))))
    return spec


def test_oracle_so2_lm_trace():
    g = load("so2_kat")
    spec = _so2_spec(g)
    np.testing.assert_allclose(nls.error_metric(spec, [v["value"] for v in spec["vars"]]), g["lm_err0"], rtol=1e-12)
    iters = g["lm_trace_err"].shape[0]
    out = nls.optimize(spec, method="lm", max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0, damping=1e-2, adaptive_damping=True,
                       ellipsoidal_damping=True)
    np.testing.assert_allclose(out["err_history"][:, 1:].T, g["lm_trace_err"], rtol=1e-9)
    for it in range(iters):
        np.testing.assert_allclose(out["trace"][it]["delta"], g["lm_trace_delta"][it], rtol=1e-6, atol=1e-10)
    np.testing.assert_allclose(np.stack(out["values"], 0), g["lm_final"], rtol=0, atol=1e-9)


def test_package_so2_class_and_torch_route():
    g = load("so2_kat")
    T = lambda k: torch.from_numpy(np.asarray(g[k]))
    X, Y, Z = th.SO2(tensor=T("X"), name="x"), th.SO2(tensor=T("Y"), name="y"), th.SO2(tensor=T("Z"), name="z")
    assert X.dof() == 1 and th.SO2.KIND == 4 and th.SO2().tensor.tolist() == [[1.0, 0.0]]
    np.testing.assert_allclose(th.SO2(theta=T("theta")).tensor.numpy(), g["X"], atol=1e-15)
    np.testing.assert_allclose(th.SO2.exp_map(T("theta")).tensor.numpy(), g["X"], atol=1e-15)
    jl = []
    np.testing.assert_allclose(X.log_map(jl).numpy(), g["log_X"], atol=1e-15)
    assert jl[0].shape == (g["X"].shape[0], 1, 1) and (jl[0] == 1).all()
    np.testing.assert_allclose(X.compose(Y).tensor.numpy(), g["compose"], atol=1e-15)
    np.testing.assert_allclose(X.inverse().tensor.numpy(), g["inverse"], atol=0)
    np.testing.assert_allclose(X.adjoint().numpy(), g["adjoint"])
    np.testing.assert_allclose(X.local(Y).numpy(), g["local"], atol=1e-15)
    np.testing.assert_allclose(X.retract(T("retract_delta")).tensor.numpy(), g["retract"], atol=1e-15)
    np.testing.assert_allclose(th.SO2.project_tensor(X.tensor, T("proj_in")).numpy(), g["proj_out"], atol=1e-15)
    np.testing.assert_allclose(X.to_matrix().numpy()[:, :, 0], g["X"], atol=0)
    p = torch.randn(g["X"].shape[0], 2, dtype=torch.float64)
    np.testing.assert_allclose(X.unrotate(X.rotate(p)).tensor.numpy(), p.numpy(), atol=1e-14)
    # cost functions: generic route (no CUDA schema), Jacobians by vmap(jacrev) + tangent projection
    cf = th.Between(X, Y, Z, th.ScaleCostWeight(torch.tensor(0.7, dtype=torch.float64)))
    assert cf.schema()[0] is None
    (J0, J1), e = cf.generic_jacobians_error([X.tensor, Y.tensor])
    np.testing.assert_allclose(e.numpy(), g["between_e"], atol=1e-14)
    np.testing.assert_allclose(J0.numpy(), g["between_J0"], atol=1e-12)
    np.testing.assert_allclose(J1.numpy(), g["between_J1"], atol=1e-12)
    cl = th.Difference(X, Z, th.ScaleCostWeight(torch.tensor(1.3, dtype=torch.float64)))
    assert cl.schema()[0] is None
    (Jl,), el = cl.generic_jacobians_error([X.tensor])
    np.testing.assert_allclose(el.numpy(), g["local_e"], atol=1e-14)
    np.testing.assert_allclose(Jl.numpy(), g["local_J"], atol=1e-12)
    np.testing.assert_allclose(cl.generic_error([X.tensor]).numpy(), g["local_e"], atol=1e-14)
    # differentiable retraction used on the tape of the backward modes
    from theseus_b200 import lie_torch
    np.testing.assert_allclose(lie_torch.retract(4, X.tensor, T("retract_delta")).numpy(), g["retract"], atol=1e-15)
    r = th.rand_so2(7, dtype=torch.float64)
    np.testing.assert_allclose((r.tensor ** 2).sum(1).numpy(), 1.0, atol=1e-14)
