"""The torch restatements that put the fused-kernel cost functions on the autograd tape (theseus_b200/lie_torch.py,
core.CostFunction._torch_error): pure torch, so they are checked on the CPU -- closed forms against the oracle, Jacobians
(vmap(jacrev) + tangent-space projection) against the reference's analytic ones (tests/golden/costs_kat.npz, so3_kat, se2_kat)."""
import numpy as np
import torch

import theseus_b200 as th
from theseus_b200 import lie_torch as lt
from oracle import lie
from helpers import load


def test_closed_forms_match_oracle():
    rng = np.random.default_rng(4)
    for scale in (1.0, 3.0, 1e-4, 1e-9):
        xi = rng.standard_normal((7, 6)) * scale
        T, U = lie.se3_exp(xi), lie.se3_exp(rng.standard_normal((7, 6)))
        tT, tU = torch.from_numpy(T), torch.from_numpy(U)
        np.testing.assert_allclose(lt.se3_exp(torch.from_numpy(xi)).numpy(), T, rtol=0, atol=1e-15)
        np.testing.assert_allclose(lt.se3_log(tT).numpy(), lie.se3_log(T), rtol=0, atol=1e-14)
        np.testing.assert_allclose(lt.local(0, tT, tU).numpy(), lie.se3_log(lie.se3_compose(lie.se3_inverse(T), U)), rtol=0, atol=1e-14)
        np.testing.assert_allclose(lt.retract(0, tT, torch.from_numpy(xi)).numpy(), lie.se3_retract(T, xi), rtol=0, atol=1e-14)
        R = lie.so3_exp(xi[:, 3:])
        np.testing.assert_allclose(lt.so3_log(torch.from_numpy(R))[0].numpy(), lie.so3_log(R), rtol=0, atol=1e-14)
        x2 = rng.standard_normal((7, 3)) * scale
        T2, U2 = lie.se2_exp(x2), lie.se2_exp(rng.standard_normal((7, 3)))
        np.testing.assert_allclose(lt.se2_exp(torch.from_numpy(x2)).numpy(), T2, rtol=0, atol=1e-15)
        np.testing.assert_allclose(lt.local(3, torch.from_numpy(T2), torch.from_numpy(U2)).numpy(),
                                   lie.se2_log(lie.se2_compose(lie.se2_inverse(T2), U2)), rtol=0, atol=1e-14)
    w = np.array([[np.pi - 1e-9, 0, 0], [0, 0, np.pi - 1e-8], [1e-9 - np.pi, 0, 0]])   # the near-pi branch
    np.testing.assert_allclose(lt.so3_log(torch.from_numpy(lie.so3_exp(w)))[0].numpy(), lie.so3_log(lie.so3_exp(w)), rtol=0, atol=1e-14)


def test_between_and_local_jacobians_match_reference_analytic_ones():
    g = load("costs_kat")
    P = lambda k: torch.from_numpy(g["f64_" + k])
    w = th.DiagonalCostWeight(P("w"))
    cf = th.Between(th.SE3(tensor=P("X0")), th.SE3(tensor=P("X1")), th.SE3(tensor=P("Z")), w)
    J, e = cf.generic_jacobians_error([P("X0"), P("X1")])
    np.testing.assert_allclose(e.numpy(), g["f64_between_e"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(J[0].numpy(), g["f64_between_J0"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(J[1].numpy(), g["f64_between_J1"], rtol=1e-9, atol=1e-10)
    lc = th.Difference(th.SE3(tensor=P("X0")), th.SE3(tensor=P("Z")), th.ScaleCostWeight(torch.tensor(0.37, dtype=torch.float64)))
    J, e = lc.generic_jacobians_error([P("X0")])
    np.testing.assert_allclose(e.numpy(), g["f64_local_e"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(J[0].numpy(), g["f64_local_J"], rtol=1e-9, atol=1e-10)
    for name, cls, lw in (("so3_kat", th.SO3, 1.3), ("se2_kat", th.SE2, 0.7)):
        g = load(name)
        X0 = torch.from_numpy(g["X0"] if name == "so3_kat" else g["exp"])
        X1 = torch.from_numpy(g["X1"] if name == "so3_kat" else g["other"])
        Z = torch.from_numpy(g["Z"])
        w = th.DiagonalCostWeight(torch.from_numpy(g["w"]))
        J, e = th.Between(cls(tensor=X0), cls(tensor=X1), cls(tensor=Z), w).generic_jacobians_error([X0, X1])
        np.testing.assert_allclose(e.numpy(), g["between_e"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(J[0].numpy(), g["between_J0"], rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(J[1].numpy(), g["between_J1"], rtol=1e-6, atol=1e-8)
        J, e = th.Difference(cls(tensor=X0), cls(tensor=Z), th.ScaleCostWeight(torch.tensor(lw, dtype=torch.float64))).generic_jacobians_error([X0])
        np.testing.assert_allclose(e.numpy(), g["local_e"], rtol=1e-10, atol=1e-12)
        # autograd differentiates the closed form exactly; the reference's analytic jlog switches to truncated series below
        # d_near_zero (1e-3 for SE2), a ~1e-8 relative difference in those rows
        np.testing.assert_allclose(J[0].numpy(), g["local_J"], rtol=1e-6, atol=1e-8)


def test_moving_frame_between_matches_reference():
    """th.eb.MovingFrameBetween (4 variables, torch path) against the reference's chained analytic Jacobians (moving_frame_kat.npz)."""
    g = load("moving_frame_kat")
    for name, cls in (("se2", th.SE2), ("se3", th.SE3)):
        ins = [torch.from_numpy(x) for x in g[f"{name}_in"]]
        vs = [cls(tensor=t) for t in ins]
        cf = th.eb.MovingFrameBetween(vs[0], vs[1], vs[2], vs[3], vs[4], th.DiagonalCostWeight(torch.from_numpy(g[f"{name}_w"])))
        assert cf.dim() == vs[0].dof() and cf.num_optim_vars() == 4 and cf.schema() == (None, [])
        J, e = cf.generic_jacobians_error(ins[:4])
        np.testing.assert_allclose(e.numpy(), g[f"{name}_e"], rtol=1e-10, atol=1e-12)
        for q in range(4):
            np.testing.assert_allclose(J[q].numpy(), g[f"{name}_J"][q], rtol=1e-8, atol=1e-10)


def test_quasi_static_pushing_planar_matches_reference():
    g = load("moving_frame_kat")
    ins = [torch.from_numpy(x) for x in g["qsp_in"]]
    vs = [th.SE2(tensor=t) for t in ins]
    cf = th.eb.QuasiStaticPushingPlanar(vs[0], vs[1], vs[2], vs[3], th.Variable(torch.from_numpy(g["qsp_c2"])),
                                        th.DiagonalCostWeight(torch.from_numpy(g["qsp_w"])))
    assert cf.dim() == 3 and cf.num_optim_vars() == 4
    J, e = cf.generic_jacobians_error(ins)
    np.testing.assert_allclose(e.numpy(), g["qsp_e"], rtol=1e-11, atol=1e-12)
    for q in range(4):
        np.testing.assert_allclose(J[q].numpy(), g["qsp_J"][q], rtol=1e-9, atol=1e-11)


def test_effector_object_contact_planar_matches_reference():
    g = load("moving_frame_kat")
    P = lambda k: torch.from_numpy(g["eoc_" + k])
    cf = th.eb.EffectorObjectContactPlanar(th.SE2(tensor=P("obj")), th.SE2(tensor=P("eff")), P("origin"), P("sdf"), 0.1,
                                           torch.tensor(0.05, dtype=torch.float64), th.ScaleCostWeight(P("w")))
    assert cf.dim() == 1
    J, e = cf.generic_jacobians_error([P("obj"), P("eff")])
    np.testing.assert_allclose(e.numpy(), g["eoc_e"], rtol=1e-11, atol=1e-13)
    assert e[-1].item() == abs(0.0 - 0.05) * float(g["eoc_w"].ravel()[0])          # out of the grid: boundary value 0
    for q in range(2):
        np.testing.assert_allclose(J[q].numpy(), g["eoc_J"][q], rtol=1e-9, atol=1e-11)


def test_tactile_objective_error_metric_matches_reference_on_the_torch_path():
    """The planar-pushing objective of tests/golden/tactile_kat.npz (config C4's cost set) assembled with this package: the sum of the
    cost functions' torch-path errors reproduces the reference's objective.error_metric() at the initial point."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    g = load("tactile_kat")
    inputs = {k: torch.from_numpy(g[k]) for k in ("obj", "eff", "eff_meas", "mfb_meas", "c_square", "eff_radius", "sdf", "sdf_origin", "sdf_cell")}
    objective, objs, effs, leaves = G.tactile_problem(th, torch, inputs)
    assert objective.size_cost_functions() == 5 * 2 + 4 * 2 + 1 and len(objective.optim_vars) == 10
    total = 0.0
    for cf in objective.cost_functions.values():
        e = cf.generic_error([v.tensor for v in cf.optim_vars])
        total = total + 0.5 * (e ** 2).sum(dim=1)
    np.testing.assert_allclose(total.numpy(), g["err0"], rtol=1e-12)
