"""Oracle SE2 closed forms + Between/Difference on SE2 + a 2-D pose-graph LM trace vs the reference (tests/golden/se2_kat.npz)."""
import numpy as np

from oracle import lie, nls
from helpers import load, se2_pg_spec


def test_se2_group_ops():
    g = load("se2_kat")
    G = g["exp"]
    np.testing.assert_allclose(lie.se2_exp(g["tangent"]), G, rtol=1e-12, atol=1e-14)
    J, xi = lie.se2_jlog(G)
    np.testing.assert_allclose(xi, g["log"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(J, g["jlog"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(lie.se2_adjoint(G), g["adj"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(lie.se2_inverse(G), g["inv"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(lie.se2_compose(G, g["other"]), g["compose"], rtol=1e-12, atol=1e-14)


def test_se2_costs():
    g = load("se2_kat")
    jacs, e = nls.between_error_jacobians("SE2", g["exp"], g["other"], g["Z"])
    jacs, e = nls.weight_jacobians_error(("diag", g["w"]), jacs, e)
    np.testing.assert_allclose(e, g["between_e"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(jacs[0], g["between_J0"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(jacs[1], g["between_J1"], rtol=1e-9, atol=1e-10)
    jl, el = nls.local_error_jacobians("SE2", g["exp"], g["Z"])
    jl, el = nls.weight_jacobians_error(("scale", np.full((1, 1), 0.7)), jl, el)
    np.testing.assert_allclose(el, g["local_e"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(jl[0], g["local_J"], rtol=1e-9, atol=1e-10)


def test_se2_pose_graph_lm_trace():
    g = load("se2_kat")
    out = nls.optimize(se2_pg_spec(g), method="lm", max_iterations=6, abs_err_tolerance=0, rel_err_tolerance=0, damping=1e-2,
                       adaptive_damping=True, ellipsoidal_damping=True)
    np.testing.assert_allclose(out["err_history"][:, 1:].T, g["pg_trace_err"], rtol=1e-8)
    for it in range(3):
        dref = g["pg_trace_delta"][it]
        rel = np.linalg.norm(out["trace"][it]["delta"] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5
