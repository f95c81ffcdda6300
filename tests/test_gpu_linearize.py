"""Fused linearize / Gram / error / retract kernels vs the reference's own numbers (golden) and the oracle."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from oracle import nls
from helpers import load, pgo_spec, pgo_objective

pytestmark = pytest.mark.gpu


def test_sparse_linearization_matches_reference():
    g = load("pgo_small_lm")
    objective, poses = pgo_objective(th, g)
    lin = th.SparseLinearization(objective)
    # integer structure: bit-exact
    assert np.array_equal(lin.A_row_ptr, g["A_row_ptr"]) and np.array_equal(lin.A_col_ind, g["A_col_ind"])
    lin.linearize()
    np.testing.assert_allclose(lin.A_val.cpu().numpy(), g["A_val0"], rtol=1e-9, atol=2e-9)
    np.testing.assert_allclose(lin.b.cpu().numpy(), g["b0"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(lin.Atb.squeeze(2).cpu().numpy(), g["Atb0_sparse"], rtol=1e-9, atol=1e-8)


def test_dense_linearization_matches_reference():
    g = load("pgo_small_lm")
    objective, poses = pgo_objective(th, g)
    lin = th.DenseLinearization(objective)
    lin.linearize()
    np.testing.assert_allclose(lin.b.cpu().numpy(), g["trace_b"][0], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(lin.AtA.cpu().numpy(), g["trace_AtA"][0], rtol=1e-9, atol=1e-7)
    np.testing.assert_allclose(lin.Atb.squeeze(2).cpu().numpy(), g["trace_Atb"][0], rtol=1e-9, atol=1e-8)
    # dense A scattered from the CSR values equals the oracle's dense A
    spec = pgo_spec(g)
    A, _, _, _ = nls.linearize_dense(spec, [v["value"] for v in spec["vars"]])
    np.testing.assert_allclose(lin.A.cpu().numpy(), A, rtol=1e-9, atol=2e-9)
    d = lin.diagonal_scaling(torch.ones_like(lin.Atb.squeeze(2)))
    np.testing.assert_allclose(d.cpu().numpy(), g["trace_AtA_diag"][0], rtol=1e-9, atol=1e-8)


def test_dense_ata_exact_over_repeated_linearizations():
    """The zero background of AtA is written once per buffer; every later linearization must still give A^T A exactly
    (pattern blocks are overwritten, nothing accumulates), also after the variables moved."""
    g = load("pgo_small_lm")
    objective, poses = pgo_objective(th, g)
    lin = th.DenseLinearization(objective)
    for rep in range(3):
        lin.linearize()
        A = lin.A
        ref = A.transpose(1, 2) @ A
        np.testing.assert_allclose(lin.AtA.cpu().numpy(), ref.cpu().numpy(), rtol=1e-11, atol=1e-9)
        for p in poses[1:4]:
            p.update(th.SE3.exp_map(0.05 * torch.randn(p.shape[0], 6, dtype=p.dtype, device=p.device)).compose(p).tensor)


def test_error_metric_and_retract_vs_oracle():
    g = load("pgo64_lm")
    objective, poses = pgo_objective(th, g)
    spec = pgo_spec(g)
    values = [v["value"] for v in spec["vars"]]
    err = objective.error_metric()
    np.testing.assert_allclose(err.cpu().numpy(), nls.error_metric(spec, values), rtol=1e-12)
    np.testing.assert_allclose(objective.error().cpu().numpy(), nls.error_vector(spec, values), rtol=1e-9, atol=1e-11)
    rng = np.random.default_rng(3)
    B, n = values[0].shape[0], 6 * len(values)
    delta = rng.standard_normal((B, n)) * 0.1
    eng = objective.engine()
    eng.adopt_optim_vars()
    tmp = [p.copy(new_name=p.name) for p in poses]
    ignore = torch.tensor([False, True], device="cuda")
    objective.retract_vars_sequence(torch.from_numpy(delta).cuda(), tmp, ignore_mask=ignore)
    want = nls.retract(spec, values, delta, ignore_mask=np.array([False, True]))
    got = np.stack([t.tensor.cpu().numpy() for t in tmp], 0)
    np.testing.assert_allclose(got, np.stack(want, 0), rtol=1e-10, atol=1e-12)
    # batch item 1 untouched bit for bit
    assert np.array_equal(got[:, 1], np.stack(values, 0)[:, 1])


def test_zero_weight_is_masked():
    g = load("pgo_small_lm")
    objective, poses = pgo_objective(th, g)
    first = next(iter(objective.cost_functions.values()))
    first.weight.diagonal.tensor = torch.zeros_like(first.weight.diagonal.tensor)
    first.measurement.tensor = torch.full_like(first.measurement.tensor, float("nan"))  # masked rows never see the NaNs
    lin = th.SparseLinearization(objective)
    lin.linearize()
    assert torch.isfinite(lin.A_val).all() and torch.isfinite(lin.b).all()
    assert (lin.b[:, :6] == 0).all()
