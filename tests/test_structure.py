"""Host logic of the product (no GPU): theseus_b200/structure.py reproduces the reference's integer structure
bit for bit (A_row_ptr, A_col_ind, var_start_cols from SparseLinearization of the real reference, stored in the
golden fixtures), and the Gram gather plan is consistent with A^T A on random values."""
import numpy as np
import pytest

from theseus_b200.structure import build_structure, build_gram_plan, ata_block_structure
from oracle import nls
from helpers import load, pgo_spec


def _structure_of(g):
    N = g["poses0"].shape[0]
    costs = [(6, (int(i), int(j))) for i, j in g["edges"]] + [(6, (0,))]
    return build_structure([6] * N, costs)


@pytest.mark.parametrize("name", ["pgo_small_lm", "pgo_small_gn", "pgo_small_lm_sph", "pgo64_lm"])
def test_csr_bit_exact_vs_reference(name):
    g = load(name)
    S = _structure_of(g)
    assert S.A_row_ptr.dtype == np.int64 and S.A_col_ind.dtype == np.int64
    assert np.array_equal(S.A_row_ptr, g["A_row_ptr"])
    assert np.array_equal(S.A_col_ind, g["A_col_ind"])
    assert np.array_equal(S.var_start_cols, g["var_start_cols"])


def test_block_pointers_with_reversed_edge():
    # an edge (j, i) with j > i: the reference sorts the column slices, so var slot 0 lands in the 2nd block
    S = build_structure([6, 6, 6], [(6, (2, 0)), (6, (0, 1))])
    assert list(S.block_pointers[0]) == [6, 0]
    assert list(S.block_pointers[1]) == [0, 6]
    assert list(S.A_col_ind[:12]) == list(range(0, 6)) + list(range(12, 18))


def test_mixed_dims_and_ata_blocks():
    S = build_structure([6, 3, 2], [(2, (0, 1)), (3, (1,)), (4, (2, 0))])
    ps, ptrs, inds = ata_block_structure(S)
    assert list(ps) == [6, 3, 2]
    assert list(ptrs) == [0, 3, 5, 7]
    assert list(inds) == [0, 1, 2, 0, 1, 0, 2]
    spec_like = nls.sparse_structure  # same routine in the oracle must agree
    assert S.num_rows == 9 and S.num_cols == 11


def test_gram_plan_equals_AtA():
    rng = np.random.default_rng(0)
    S = build_structure([6, 3, 2, 6], [(2, (0, 1)), (3, (1,)), (4, (2, 0)), (6, (3, 0)), (6, (0, 3))])
    plan = build_gram_plan(S)
    A_val = rng.standard_normal(S.nnz)
    b = rng.standard_normal(S.num_rows)
    A = np.zeros((S.num_rows, S.num_cols))
    for r in range(S.num_rows):
        A[r, S.A_col_ind[S.A_row_ptr[r]:S.A_row_ptr[r + 1]]] = A_val[S.A_row_ptr[r]:S.A_row_ptr[r + 1]]
    n = S.num_cols
    out = np.zeros(n * n)
    for e in range(plan["ent_blk"].shape[0]):
        blk, p, q = plan["ent_blk"][e], plan["ent_p"][e], plan["ent_q"][e]
        acc = 0.0
        for c in range(plan["blk_cptr"][blk], plan["blk_cptr"][blk + 1]):
            for r in range(plan["c_rows"][c]):
                base = plan["c_off"][c] + r * plan["c_stride"][c]
                acc += A_val[base + plan["c_bpa"][c] + p] * A_val[base + plan["c_bpb"][c] + q]
        out[plan["blk_out"][blk] + p * plan["blk_ld"][blk] + q] = acc
        if plan["blk_mirror"][blk] >= 0:
            out[plan["blk_mirror"][blk] + q * plan["blk_ld"][blk] + p] = acc
    np.testing.assert_allclose(out.reshape(n, n), A.T @ A, rtol=1e-12, atol=1e-12)
    atb = np.zeros(n)
    for col in range(n):
        for c in range(plan["col_cptr"][col], plan["col_cptr"][col + 1]):
            for r in range(plan["cc_rows"][c]):
                atb[col] += A_val[plan["cc_off"][c] + r * plan["cc_stride"][c]] * b[plan["cc_row0"][c] + r]
    np.testing.assert_allclose(atb, A.T @ b, rtol=1e-12, atol=1e-12)
