"""Robust losses (theseus/core/robust_loss.py, robust_cost_function.py): the oracle's restatement and the package's torch route
(RobustCostFunction.generic_jacobians_error / generic_error: Hinge, Geman-McClure, flatten_dims and wrapped Vector differences are
served by the engine's generic route with exactly these functions) against values computed by the reference
(tests/golden/make_golden.py:make_robust -> robust_kat.npz)."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from helpers import load
from oracle import nls

LOSSES = dict(welsch=th.WelschLoss, huber=th.HuberLoss, hinge=th.HingeLoss, geman=th.GemanMcClureLoss)


def _oracle_inner(g, cname):
    if cname == "between":
        jacs, e = nls.between_error_jacobians("SE3", g["X0"], g["X1"], g["Z"])
        return nls.weight_jacobians_error(("diag", g["w"]), jacs, e)
    e = g["V"] - g["Vt"]                                                   # Vector.local (geometry/vector.py): x - target, J = I
    jacs = [np.broadcast_to(np.eye(e.shape[1]), e.shape + (e.shape[1],)).copy()]
    return nls.weight_jacobians_error(("scale", np.full((1, 1), float(g["wv"]))), jacs, e)


@pytest.mark.parametrize("cname", ["between", "vecdiff"])
@pytest.mark.parametrize("flat", [False, True])
@pytest.mark.parametrize("name", list(LOSSES))
def test_oracle_robust_apply_matches_reference(name, flat, cname):
    g = load("robust_kat")
    jacs, e = _oracle_inner(g, cname)
    robust = (name, g["log_radius"]) + ((g["mu"],) if name == "geman" else ())
    rj, re = nls.robust_apply(robust, jacs, e, flatten_dims=flat)
    key = f"{name}_{'flat' if flat else 'full'}_{cname}"
    np.testing.assert_allclose(re, g[key + "_e"], rtol=1e-10, atol=1e-13)
    for k in range(len(rj)):
        np.testing.assert_allclose(rj[k], g[key + "_J"][k], rtol=1e-9, atol=1e-12)
    _, val = nls.robust_apply(robust, None, e, flatten_dims=flat)
    np.testing.assert_allclose(val, g[key + "_val"], rtol=1e-10, atol=1e-13)
    if name in ("huber", "hinge") and not flat:   # both branches of the piecewise losses are exercised
        x = (e ** 2).sum(-1)
        assert (x > np.exp(g["log_radius"])).any() and (x < np.exp(g["log_radius"])).any()


def _package_cost(g, name, flat, cname):
    d = torch.float64
    T = lambda k: torch.from_numpy(np.asarray(g[k])).to(d)
    if cname == "between":
        inner = th.Between(th.SE3(tensor=T("X0"), name="x0"), th.SE3(tensor=T("X1"), name="x1"), th.SE3(tensor=T("Z"), name="z"),
                           th.DiagonalCostWeight(T("w")))
    else:
        inner = th.Difference(th.Vector(tensor=T("V"), name="v"), th.Vector(tensor=T("Vt"), name="vt"),
                              th.ScaleCostWeight(torch.tensor(float(g["wv"]), dtype=d)))
    lr = th.Vector(tensor=T("log_radius"), name="lr")
    if name == "geman":
        return th.GNCRobustCostFunction(inner, LOSSES[name], lr, th.Vector(tensor=T("mu"), name="mu"), flatten_dims=flat)
    return th.RobustCostFunction(inner, LOSSES[name], lr, flatten_dims=flat)


@pytest.mark.parametrize("cname", ["between", "vecdiff"])
@pytest.mark.parametrize("flat", [False, True])
@pytest.mark.parametrize("name", list(LOSSES))
def test_package_torch_route_matches_reference(name, flat, cname):
    g = load("robust_kat")
    cf = _package_cost(g, name, flat, cname)
    ts = [v.tensor for v in cf.optim_vars]
    jacs, e = cf.generic_jacobians_error(ts)
    key = f"{name}_{'flat' if flat else 'full'}_{cname}"
    np.testing.assert_allclose(e.numpy(), g[key + "_e"], rtol=1e-9, atol=1e-12)
    for k in range(len(jacs)):
        np.testing.assert_allclose(jacs[k].numpy(), g[key + "_J"][k], rtol=1e-7, atol=1e-9)   # vmap(jacrev) vs the analytic Jacobians
    np.testing.assert_allclose(cf.generic_error(ts).numpy(), g[key + "_val"], rtol=1e-9, atol=1e-12)
    # routing: fused kernels only for Welsch / Huber around a cost with a CUDA schema, without flatten_dims
    kind, aux = cf.schema()
    fused = name in ("welsch", "huber") and not flat and cname == "between"
    assert (kind is not None) == fused
    assert cf.log_loss_radius in cf.aux_vars and (name != "geman" or cf.gnc_control_val in cf.aux_vars)


def test_loss_classes_have_the_reference_interface():
    x = torch.tensor([[0.1], [2.0]], dtype=torch.float64)
    lr = torch.tensor([[0.0]], dtype=torch.float64)
    for cls in (th.WelschLoss, th.HuberLoss, th.HingeLoss):
        assert cls.evaluate(x, lr).shape == x.shape and cls.linearize(x, lr).shape == x.shape
    mu = torch.tensor([[1.0]], dtype=torch.float64)
    np.testing.assert_allclose(th.GemanMcClureLoss.evaluate(x, lr, mu).numpy(), (x / (1 + x)).numpy(), rtol=1e-14)
    with pytest.raises(RuntimeError):
        th.GNCRobustCostFunction(None, th.HuberLoss, None, None)
    assert issubclass(th.GemanMcClureLoss, th.GNCRobustLoss) and issubclass(th.GNCRobustLoss, th.RobustLoss)
