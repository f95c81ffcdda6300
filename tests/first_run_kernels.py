"""NOT COLLECTED by the suite (no test_ prefix): the first-run tests of the kernels that are NEW (dense root, chain-piece substitutions,
tiled updates).  tests/test_gpu_zz_first_run.py runs this file in CHILD processes (one per group), so that a fault in a kernel that has
never executed on a device -- a sticky CUDA error, a crash at interpreter exit -- cannot reach the process that reports the GPU-verified
suite.  Directly:  python -m pytest tests/first_run_kernels.py -m gpu   (or with THB_SIMT_EMULATION=1 on the host emulation)."""
import numpy as np
import pytest
import torch

import theseus_b200 as th
from helpers import load, decisive_iterations

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("layout", ["lane_root"])
def test_c5_full_size_sparse_lm_trace(layout):
    """Config C5's pose graph at full size (2 500 poses, n = 15 000), one batch item: the block-sparse solver's LM trace against the
    reference's dense-solver trace (tests/golden/pgo_c5_lm.npz, generated on the CPU by make_golden.py c5).  Same parked status."""
    from helpers import pgo_objective, lm_kwargs_of
    g = load("pgo_c5_lm")
    method, iters, kw = lm_kwargs_of(g)
    objective, poses = pgo_objective(th, g)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                                max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0,
                                linear_solver_kwargs=dict(layout=layout))
    errs, deltas = [], []

    def cb(optimizer, info, delta, it):
        errs.append(info.last_err.cpu().numpy().copy()); deltas.append(delta.cpu().numpy().copy())
    with torch.no_grad():
        np.testing.assert_allclose(objective.error_metric().cpu().numpy(), g["err0"], rtol=1e-10)
        opt.optimize(end_iter_callback=cb, **kw)
    np.testing.assert_allclose(np.stack(errs, 0), g["trace_err"], rtol=1e-7)
    for it in range(decisive_iterations(g["err0"], g["trace_err"])):
        rel = np.linalg.norm(deltas[it] - g["trace_delta"][it], axis=1) / np.linalg.norm(g["trace_delta"][it], axis=1)
        assert rel.max() < 1e-5, (it, rel)


def test_lane_root_layout_matches_lane_layout():
    """Opt-in layout='lane_root' (dense DMMA factorisation of the top chain of the elimination tree, sparse.root_split) against the plain
    lane layout and the dense residual, on a ring-with-chords structure whose minimum-degree order ends in a dense separator chain.
    The host half (work lists) is verified on the CPU: tests/test_sparse_symbolic.py::test_root_split_solves_system."""
    from theseus_b200.structure import build_structure
    from test_gpu_sparse_solver import _dense_system
    rng = np.random.default_rng(5)
    N, B = 60, 70
    costs = [(3, [i, (i + 1) % N]) for i in range(N)] + [(3, [i, (i + 7) % N]) for i in range(N)] + [(6, [i]) for i in range(N)]
    costs = [(d, sorted(vs)) for d, vs in costs]
    S = build_structure([6] * N, costs)
    A_val = torch.from_numpy(rng.standard_normal((B, S.nnz))).cuda()
    b = torch.from_numpy(rng.standard_normal((B, S.num_rows))).cuda()
    alpha = torch.from_numpy(rng.random(B) * 0.1).cuda()
    xs = {}
    for layout in ("lane_root", "lane"):
        solver = th.BaspachoSparseSolver.from_structure(S, layout=layout)
        solver.linearization.A_val, solver.linearization.b = A_val, b
        xs[layout] = (solver.solve(damping=alpha, ellipsoidal_damping=True, damping_eps=1e-6).cpu().numpy(), solver.solve().cpu().numpy())
        if layout == "lane_root":
            assert solver._root is not None and solver._dev["nt"] >= 48
    AtA, Atb = _dense_system(S, A_val, b)
    idx = np.arange(S.num_cols)
    for k, (mul, add) in enumerate(((1 + alpha.cpu().numpy()[:, None], 1e-6), (1.0, 0.0))):
        M = AtA.copy()
        M[:, idx, idx] = M[:, idx, idx] * mul + add
        res = np.einsum("bij,bj->bi", M, xs["lane_root"][k]) - Atb
        assert np.abs(res).max() < 1e-10 * np.abs(M).sum(axis=2).max() * max(1.0, np.abs(xs["lane_root"][k]).max())
        assert np.abs(xs["lane_root"][k] - xs["lane"][k]).max() < 1e-11 * np.linalg.cond(M).max() * max(1.0, np.abs(xs["lane"][k]).max())


@pytest.mark.parametrize("layout", ["lane", "lane_root", "lane_tiled_root"])
@pytest.mark.parametrize("B", [32, 70])
def test_supernodal_substitutions_match_per_column_substitutions(layout, B):
    """supernodal_solve=True (chain-piece forward / backward kernels, thb_sparse_lane.cu:lane_piece_forward_kernel / _backward_kernel,
    lists sparse.piece_solve_lists) against the per-column substitution kernels on the same factor, and the dense residual.  The schedule
    is verified on the CPU: tests/test_sparse_symbolic.py::test_piece_solve_schedule_solves_system."""
    from theseus_b200.structure import build_structure
    from test_gpu_sparse_solver import _dense_system
    rng = np.random.default_rng(19 + B)
    N = 60
    costs = [(3, [i, (i + 1) % N]) for i in range(N)] + [(3, [i, (i + 7) % N]) for i in range(N)] + [(6, [i]) for i in range(N)]
    costs = [(d, sorted(vs)) for d, vs in costs]
    S = build_structure([6] * N, costs)
    A_val = torch.from_numpy(rng.standard_normal((B, S.nnz))).cuda()
    b = torch.from_numpy(rng.standard_normal((B, S.num_rows))).cuda()
    alpha = torch.from_numpy(rng.random(B) * 0.1).cuda()
    xs = {}
    for sn in (True, False):
        solver = th.BaspachoSparseSolver.from_structure(S, layout=layout, supernodal_solve=sn)
        solver.linearization.A_val, solver.linearization.b = A_val, b
        xs[sn] = solver.solve(damping=alpha, ellipsoidal_damping=True, damping_eps=1e-6).cpu().numpy()
        assert ("pieces" in solver._dev) == sn
    AtA, Atb = _dense_system(S, A_val, b)
    idx = np.arange(S.num_cols)
    M = AtA.copy()
    M[:, idx, idx] = M[:, idx, idx] * (1 + alpha.cpu().numpy()[:, None]) + 1e-6
    res = np.einsum("bij,bj->bi", M, xs[True]) - Atb
    assert np.abs(res).max() < 1e-10 * np.abs(M).sum(axis=2).max() * max(1.0, np.abs(xs[True]).max())
    assert np.abs(xs[True] - xs[False]).max() < 1e-11 * np.linalg.cond(M).max() * max(1.0, np.abs(xs[False]).max())


@pytest.mark.parametrize("tiled", ["lane_tiled", "lane_tiled_root"])
@pytest.mark.parametrize("B", [32, 70])
def test_lane_tiled_layout_matches_lane_layout(B, tiled):
    """Opt-in layouts 'lane_tiled' / 'lane_tiled_root' (the latter: + dense root, its assembly as one tile launch) (external updates of chain pieces as 4x4 tiles with the source blocks staged in shared memory,
    thb_sparse_lane.cu:lane_tile_update_kernel) against the plain lane layout and the dense residual; same ring-with-chords structure
    (its elimination tree has chains of every width up to the dense separator).  The host half (flat tile arrays) is verified on the
    CPU: tests/test_sparse_symbolic.py::test_tiled_lane_lists_solve_system."""
    from theseus_b200.structure import build_structure
    from test_gpu_sparse_solver import _dense_system
    rng = np.random.default_rng(11 + B)
    N = 60
    costs = [(3, [i, (i + 1) % N]) for i in range(N)] + [(3, [i, (i + 7) % N]) for i in range(N)] + [(6, [i]) for i in range(N)]
    costs = [(d, sorted(vs)) for d, vs in costs]
    S = build_structure([6] * N, costs)
    A_val = torch.from_numpy(rng.standard_normal((B, S.nnz))).cuda()
    b = torch.from_numpy(rng.standard_normal((B, S.num_rows))).cuda()
    alpha = torch.from_numpy(rng.random(B) * 0.1).cuda()
    xs = {}
    for layout in (tiled, "lane"):
        solver = th.BaspachoSparseSolver.from_structure(S, layout=layout)
        solver.linearization.A_val, solver.linearization.b = A_val, b
        xs[layout] = (solver.solve(damping=alpha, ellipsoidal_damping=True, damping_eps=1e-6).cpu().numpy(), solver.solve().cpu().numpy())
        if layout == tiled:
            assert solver._tiles[1]["tile_tgt"].shape[0] > 0
    xs["lane_tiled"] = xs[tiled]
    AtA, Atb = _dense_system(S, A_val, b)
    idx = np.arange(S.num_cols)
    for k, (mul, add) in enumerate(((1 + alpha.cpu().numpy()[:, None], 1e-6), (1.0, 0.0))):
        M = AtA.copy()
        M[:, idx, idx] = M[:, idx, idx] * mul + add
        res = np.einsum("bij,bj->bi", M, xs["lane_tiled"][k]) - Atb
        assert np.abs(res).max() < 1e-10 * np.abs(M).sum(axis=2).max() * max(1.0, np.abs(xs["lane_tiled"][k]).max())
        assert np.abs(xs["lane_tiled"][k] - xs["lane"][k]).max() < 1e-11 * np.linalg.cond(M).max() * max(1.0, np.abs(xs["lane"][k]).max())
