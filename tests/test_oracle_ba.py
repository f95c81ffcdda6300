"""Oracle vs the reference on the bundle-adjustment fixture (Reprojection + Difference on SE3 / Point3): CSR structure
bit-exact, linearization, and the complete LM trace."""
import numpy as np
import pytest

from oracle import nls
from helpers import load, ba_spec, lm_kwargs_of, decisive_iterations


@pytest.mark.parametrize("name", ["ba_small_lm", "ba_small_huber"])
def test_ba_structure_and_linearization(name):
    g = load(name)
    spec = ba_spec(g)
    st = nls.sparse_structure(spec)
    assert np.array_equal(st["A_row_ptr"], g["A_row_ptr"]) and np.array_equal(st["A_col_ind"], g["A_col_ind"])
    A_val, b = nls.linearize_sparse(spec, [v["value"] for v in spec["vars"]], st)
    np.testing.assert_allclose(A_val, g["A_val0"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(b, g["b0"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", ["ba_small_lm", "ba_small_huber"])
def test_ba_lm_trace(name):
    g = load(name)
    method, iters, kw = lm_kwargs_of(g)
    spec = ba_spec(g)
    out = nls.optimize(spec, method=method, max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0, **kw)
    np.testing.assert_allclose(out["err_history"][:, 1:].T, g["trace_err"], rtol=1e-8)
    err0 = nls.error_metric(spec, [v["value"] for v in spec["vars"]])
    k = decisive_iterations(err0, g["trace_err"])
    assert k >= 3
    for it in range(k):
        dref = g["trace_delta"][it]
        rel = np.linalg.norm(out["trace"][it]["delta"] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5, (it, rel)
