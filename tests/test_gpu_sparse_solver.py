"""Batched block-sparse Cholesky (BaspachoSparseSolver mirror) on the GPU, tested the way the reference tests its sparse
solvers: random block-sparse systems, residual of (AtA + damping) x = Atb  (tests/theseus_tests/extlib/test_baspacho.py:16-150,
optimizer/linear/test_baspacho_sparse_solver.py:15-92), plus end-to-end LM against the reference's dense-solver traces
(the reference pins Dense == BaSpaCho == LU-CUDA to 1e-10 in test_pgo_benchmark.py:34-71)."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from theseus_b200.structure import build_structure
from helpers import load, pgo_objective, lm_kwargs_of, decisive_iterations, pgo_spec
from oracle import nls

pytestmark = pytest.mark.gpu


def _random_structure(rng, num_cols, param_size_range, fill, num_rows_blocks):
    sizes = []
    while sum(sizes) < num_cols:
        if isinstance(param_size_range, list):
            sizes.append(int(rng.choice(param_size_range)))
        else:
            sizes.append(int(rng.integers(*param_size_range)))
    N = len(sizes)
    costs = []
    for _ in range(num_rows_blocks):
        vs = [v for v in range(N) if rng.random() < fill]
        if len(vs) < 1:
            vs = [int(rng.integers(N))]
        vs = vs[:2] if len(vs) > 2 else vs       # the engine's cost groups hold <= 2 variables; the solver itself is general
        costs.append((int(rng.integers(1, 4)), vs))
    for v in range(N):                            # every variable observed at least once -> AtA + damping is SPD
        costs.append((sizes[v], [v]))
    return build_structure(sizes, costs)


@pytest.mark.parametrize("B,num_cols,psr,fill", [(1, 30, (2, 6), 0.05), (32, 30, (2, 6), 0.05), (128, 70, (1, 13), 0.02), (32, 70, (2, 6), 0.05)])
@pytest.mark.parametrize("ordering", ["mindeg", "natural"])
def test_random_block_sparse_systems(B, num_cols, psr, fill, ordering):
    rng = np.random.default_rng(B + num_cols)
    S = _random_structure(rng, num_cols, psr, fill, num_rows_blocks=3 * num_cols)
    solver = th.BaspachoSparseSolver.from_structure(S, ordering=ordering)
    A_val = torch.from_numpy(rng.standard_normal((B, S.nnz))).cuda()
    b = torch.from_numpy(rng.standard_normal((B, S.num_rows))).cuda()
    solver.linearization.A_val, solver.linearization.b = A_val, b
    alpha = torch.from_numpy(rng.random(B) * 0.1).cuda()
    for ell, damping in ((True, alpha), (False, alpha), (False, 0.37), (None, None)):
        x = solver.solve(damping=damping, ellipsoidal_damping=bool(ell), damping_eps=1e-6) if damping is not None else solver.solve()
        # dense reconstruction of the system, as the reference's test does with scipy
        A = np.zeros((B, S.num_rows, S.num_cols))
        for r in range(S.num_rows):
            A[:, r, S.A_col_ind[S.A_row_ptr[r]:S.A_row_ptr[r + 1]]] = A_val.cpu().numpy()[:, S.A_row_ptr[r]:S.A_row_ptr[r + 1]]
        AtA = np.einsum("bri,brj->bij", A, A)
        Atb = np.einsum("bri,br->bi", A, b.cpu().numpy())
        idx = np.arange(S.num_cols)
        if damping is not None:
            dv = damping.cpu().numpy() if torch.is_tensor(damping) else np.full(B, damping)
            if ell:
                AtA[:, idx, idx] = AtA[:, idx, idx] * (1 + dv[:, None]) + 1e-6   # diag*(1+alpha)+beta (baspacho_solver.cpp:181-183)
            else:
                AtA[:, idx, idx] += dv[:, None]
        res = np.einsum("bij,bj->bi", AtA, x.cpu().numpy()) - Atb
        scale = np.abs(AtA).sum(axis=2).max() * max(1.0, np.abs(x.cpu().numpy()).max())
        assert np.abs(res).max() < 1e-10 * scale, (ell, np.abs(res).max())


def _dense_system(S, A_val, b):
    B = A_val.shape[0]
    A = np.zeros((B, S.num_rows, S.num_cols))
    for r in range(S.num_rows):
        A[:, r, S.A_col_ind[S.A_row_ptr[r]:S.A_row_ptr[r + 1]]] = A_val.cpu().numpy()[:, S.A_row_ptr[r]:S.A_row_ptr[r + 1]]
    return np.einsum("bri,brj->bij", A, A), np.einsum("bri,br->bi", A, b.cpu().numpy())


def _clique_structure(rng, sizes):
    """Every pair of variables shares a cost function: AtA is block-dense, the last blocks collect one update pair per earlier column."""
    N = len(sizes)
    costs = [(int(rng.integers(1, 4)), [i, j]) for i in range(N) for j in range(i + 1, N)] + [(sizes[v], [v]) for v in range(N)]
    return build_structure(sizes, costs)


@pytest.mark.parametrize("B,num_cols,sizes,fill", [(1, 40, [1, 2, 3, 6], 0.05), (32, 60, [3, 6], 0.05), (70, 90, [1, 2, 3, 6], 0.03),
                                                    (33, 0, [6] * 20, -1), (64, 0, [3, 6, 6, 2, 1, 6] * 4, -1)])
@pytest.mark.parametrize("ordering", ["mindeg", "natural"])
def test_lane_layout_matches_item_layout_and_residual(B, num_cols, sizes, fill, ordering):
    """The batch-interleaved kernels (thb_sparse_lane.cu) against the one-CTA-per-item kernels and the dense residual; the clique
    cases (fill -1) put >= 8 update pairs on the last blocks and so cover the split-K update kernel; B = 33, 70: ragged warps."""
    rng = np.random.default_rng(7 * B + num_cols)
    S = _clique_structure(rng, sizes) if fill < 0 else _random_structure(rng, num_cols, sizes, fill, num_rows_blocks=3 * num_cols)
    A_val = torch.from_numpy(rng.standard_normal((B, S.nnz))).cuda()
    b = torch.from_numpy(rng.standard_normal((B, S.num_rows))).cuda()
    alpha = torch.from_numpy(rng.random(B) * 0.1).cuda()
    xs = {}
    for layout in ("lane", "item"):
        solver = th.BaspachoSparseSolver.from_structure(S, ordering=ordering, layout=layout)
        assert solver.layout_for(B) == layout
        solver.linearization.A_val, solver.linearization.b = A_val, b
        xs[layout] = (solver.solve(damping=alpha, ellipsoidal_damping=True, damping_eps=1e-6).cpu().numpy(),
                      solver.solve(damping=0.37, ellipsoidal_damping=False).cpu().numpy(), solver.solve().cpu().numpy())
    if fill < 0:
        assert (solver._plan.lane["launches"][:, 0] == 3).any()   # heavy (split-K) update launches present
    AtA, Atb = _dense_system(S, A_val, b)
    idx = np.arange(S.num_cols)
    # a python-float damping goes through torch.as_tensor (fp32) in the reference too (optimizer/linear/utils.py:20)
    for k, (mul, add) in enumerate(((1 + alpha.cpu().numpy()[:, None], 1e-6), (1.0, float(np.float32(0.37))), (1.0, 0.0))):
        M = AtA.copy()
        M[:, idx, idx] = M[:, idx, idx] * mul + add
        x = xs["lane"][k]
        res = np.einsum("bij,bj->bi", M, x) - Atb
        scale = np.abs(M).sum(axis=2).max() * max(1.0, np.abs(x).max())
        assert np.abs(res).max() < 1e-10 * scale
        ref = np.linalg.solve(M, Atb[..., None])[..., 0]
        cond = np.linalg.cond(M).max()
        assert np.abs(x - ref).max() <= 1e-11 * cond * max(1.0, np.abs(ref).max())
        assert np.abs(x - xs["item"][k]).max() <= 1e-11 * cond * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("layout", ["lane", "item"])
def test_not_positive_definite_raises(layout):
    S = build_structure([2, 2], [(2, [0, 1])])   # rank-deficient: 2 rows, 4 columns, no damping
    solver = th.BaspachoSparseSolver.from_structure(S, layout=layout)
    solver.linearization.A_val = torch.ones(3, S.nnz, dtype=torch.float64, device="cuda")
    solver.linearization.b = torch.ones(3, S.num_rows, dtype=torch.float64, device="cuda")
    with pytest.raises(RuntimeError, match="positive definite"):
        solver.solve()


@pytest.mark.parametrize("layout", ["lane", "item"])
@pytest.mark.parametrize("name", ["pgo_small_lm", "pgo_small_lm_hard", "pgo32_lm_hard", "pgo64_lm"])
def test_lm_with_sparse_solver_matches_reference_traces(name, layout):
    g = load(name)
    method, iters, kw = lm_kwargs_of(g)
    objective, poses = pgo_objective(th, g)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                                max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0,
                                linear_solver_kwargs=dict(layout=layout))
    # block structure handed to the symbolic phase == the oracle's restatement of baspacho_sparse_solver.py:93-113 (bit-exact)
    ps, ptrs, inds = nls.ata_block_structure(pgo_spec(g))
    assert np.array_equal(opt.linear_solver.param_size, ps) and np.array_equal(opt.linear_solver.block_ptrs, ptrs)
    assert np.array_equal(opt.linear_solver.block_inds, inds)
    deltas, errs = [], []

    def cb(optimizer, info, delta, it):
        deltas.append(delta.cpu().numpy().copy())
        errs.append(info.last_err.cpu().numpy().copy())

    with torch.no_grad():
        info = opt.optimize(end_iter_callback=cb, **kw)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-8)
    spec = pgo_spec(g)
    err0 = nls.error_metric(spec, [v["value"] for v in spec["vars"]])
    k = decisive_iterations(err0, g["trace_err"])
    for it in range(k):
        dref = g["trace_delta"][it]
        rel = np.linalg.norm(deltas[it] - dref, axis=1) / np.linalg.norm(dref, axis=1)
        assert rel.max() < 1e-5, (it, rel)
