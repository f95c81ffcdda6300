"""Host logic of the block-sparse solver (no GPU): the symbolic plan (ordering, fill, levels, update lists, work items)
is executed by a tiny numpy interpreter of the very arrays the CUDA kernels consume, and the result is checked the way
the reference checks BaSpaCho (tests/theseus_tests/extlib/test_baspacho.py:101-114: residual of (AtA + damping) x = Atb),
including the reference's literal 12x12 system shape (params [2,3,5,2], extlib/test_baspacho_simple.py:15-112)."""
import numpy as np
import pytest

from theseus_b200.sparse import analyze, analyze_py, chain_tiles, minimum_degree_order, root_split


def run_plan_numpy(plan, M, rhs):
    """Interpret plan.arrays on one SPD matrix M (original order).  Returns x with M x = rhs."""
    A = plan.arrays
    N, dims, cs = plan.N, plan.dims, plan.col_start
    F = np.zeros(plan.data_size)
    for (i, j), t in plan.blk_index.items():
        blk = M[cs[i]:cs[i] + dims[i], cs[j]:cs[j] + dims[j]]
        F[plan.blk_off[t]:plan.blk_off[t] + blk.size] = blk.reshape(-1)
    W = np.zeros(plan.winv_size)
    nlev = len(A["u_ptr"]) - 1
    for lv in range(nlev):
        for e in range(A["u_ptr"][lv], A["u_ptr"][lv + 1]):
            tgt, di, dj, diag = A["u_tgt"][e], A["u_r"][e], A["u_c"][e], A["u_ld"][e]
            for r in range(di):
                for c in range(dj):
                    if diag and c > r:
                        continue
                    acc = F[tgt + r * dj + c]
                    for p in range(A["u_p0"][e], A["u_p1"][e]):
                        dk = A["up_k"][p]
                        a0, b0 = A["up_a"][p] + r * dk, A["up_b"][p] + c * dk
                        acc -= F[a0:a0 + dk] @ F[b0:b0 + dk]
                    F[tgt + r * dj + c] = acc
        for e in range(A["f_ptr"][lv], A["f_ptr"][lv + 1]):
            off, d, w = A["f_off"][e], A["f_dim"][e], A["f_w"][e]
            D = np.tril(F[off:off + d * d].reshape(d, d))
            D = D + np.tril(D, -1).T
            L = np.linalg.cholesky(D)
            F[off:off + d * d] = L.reshape(-1)
            W[w:w + d * d] = np.linalg.inv(L).reshape(-1)
        for e in range(A["t_ptr"][lv], A["t_ptr"][lv + 1]):
            off, r, d, w = A["t_off"][e], A["t_r"][e], A["t_dim"][e], A["t_w"][e]
            U = F[off + r * d:off + r * d + d].copy()
            Wm = W[w:w + d * d].reshape(d, d)
            F[off + r * d:off + r * d + d] = Wm @ U  # L[r][c] = sum_q U[q] W[c][q]
    # solve
    y = [rhs[cs[j]:cs[j] + dims[j]].copy() for j in range(N)]
    for lv in range(nlev):
        for e in range(A["s_ptr"][lv], A["s_ptr"][lv + 1]):
            j = A["s_col"][e]
            d = dims[j]
            s = y[j].copy()
            for p in range(A["fr_ptr"][j], A["fr_ptr"][j + 1]):
                k = A["fr_k"][p]
                dk = dims[k]
                s -= F[A["fr_off"][p]:A["fr_off"][p] + d * dk].reshape(d, dk) @ y[k]
            y[j] = W[A["winv_off"][j]:A["winv_off"][j] + d * d].reshape(d, d) @ s
    for lv in reversed(range(nlev)):
        for e in range(A["s_ptr"][lv], A["s_ptr"][lv + 1]):
            j = A["s_col"][e]
            d = dims[j]
            s = y[j].copy()
            for p in range(A["bc_ptr"][j], A["bc_ptr"][j + 1]):
                i = A["bc_i"][p]
                di = dims[i]
                s -= F[A["bc_off"][p]:A["bc_off"][p] + di * d].reshape(di, d).T @ y[i]
            y[j] = W[A["winv_off"][j]:A["winv_off"][j] + d * d].reshape(d, d).T @ s
    x = np.zeros_like(rhs)
    for j in range(N):
        x[cs[j]:cs[j] + dims[j]] = y[j]
    return x


def apply_tiles_numpy(F, tiles, b0, b1, di, dj):
    """Launch row (TU, 6, 6, b0, b1): what lane_tile_update_kernel<6,4,4> does to the factor storage."""
    TR, TC, D = 4, 4, 6
    assert di == D and dj == D
    for t in range(b0, b1):
        tg = tiles["tile_tgt"][t]
        acc = {w: F[tg[w]:tg[w] + D * D].reshape(D, D).copy() for w in range(TR * TC) if tg[w] >= 0}
        assert tiles["step_ptr"][t + 1] > tiles["step_ptr"][t]
        for st in range(tiles["step_ptr"][t], tiles["step_ptr"][t + 1]):
            src = tiles["step_src"][st]
            for w in acc:
                ro, co = src[w // TC], src[TR + w % TC]
                if ro >= 0 and co >= 0:
                    acc[w] -= F[ro:ro + D * D].reshape(D, D) @ F[co:co + D * D].reshape(D, D).T
        for w, V in acc.items():
            F[tg[w]:tg[w] + D * D] = V.reshape(-1)


def run_lane_plan_numpy(plan, M, rhs, lane=None, tiles=None):
    """Interpret plan.lane (the work lists of thb_sparse_lane.cu) in launch order: U / UH = left-looking update of a block,
    T = Cholesky of the column's diagonal block (to `diagl`, reciprocal diagonal) + triangular solve of the block, S = substitutions,
    TU (with `tiles`, sparse.tile_lane_lists) = tiled external updates: per k step every target with both sources present."""
    A, Ln = plan.arrays, (plan.lane if lane is None else lane)
    N, dims, cs = plan.N, plan.dims, plan.col_start
    F = np.zeros(plan.data_size)
    for (i, j), t in plan.blk_index.items():
        blk = M[cs[i]:cs[i] + dims[i], cs[j]:cs[j] + dims[j]]
        F[plan.blk_off[t]:plan.blk_off[t] + blk.size] = blk.reshape(-1)
    DL = np.zeros(plan.winv_size)
    seen_u = set()
    for kind, di, dj, b0, b1 in Ln["launches"]:
        if kind in (0, 3):
            for e in range(b0, b1):
                tgt = Ln["u_tgt"][e]
                assert tgt not in seen_u
                seen_u.add(tgt)
                T = F[tgt:tgt + di * dj].reshape(di, dj).copy()
                npairs = Ln["u_p1"][e] - Ln["u_p0"][e]
                assert (npairs >= 8) == (kind == 3)
                for p in range(Ln["u_p0"][e], Ln["u_p1"][e]):
                    dk = A["up_k"][p]
                    T -= F[A["up_a"][p]:A["up_a"][p] + di * dk].reshape(di, dk) @ F[A["up_b"][p]:A["up_b"][p] + dj * dk].reshape(dj, dk).T
                F[tgt:tgt + di * dj] = T.reshape(-1)
        elif kind == 4:
            apply_tiles_numpy(F, tiles, b0, b1, di, dj)
        elif kind == 1:
            diag_writes = {}
            for e in range(b0, b1):   # every item reads the PRE-factor diagonal block: collect, then write
                off, dg, dl = Ln["t_off"][e], Ln["t_diag"][e], Ln["t_dl"][e]
                D = np.tril(F[dg:dg + dj * dj].reshape(dj, dj))
                L = np.linalg.cholesky(D + np.tril(D, -1).T)
                if off == dg:
                    assert di == dj
                    Lr = L.copy()
                    Lr[np.arange(dj), np.arange(dj)] = 1.0 / np.diag(L)
                    diag_writes[dl] = Lr
                else:
                    X = np.linalg.solve(L, F[off:off + di * dj].reshape(di, dj).T).T   # X L^T = U
                    F[off:off + di * dj] = X.reshape(-1)
            for dl, Lr in diag_writes.items():
                DL[dl:dl + Lr.size] = Lr.reshape(-1)

    def diag_solve(j, s, transpose):
        d = dims[j]
        Lr = DL[A["winv_off"][j]:A["winv_off"][j] + d * d].reshape(d, d).copy()
        Lr[np.arange(d), np.arange(d)] = 1.0 / np.diag(Lr)
        return np.linalg.solve(Lr.T if transpose else Lr, s)

    y = [None] * N
    S_launches = [l for l in Ln["launches"] if l[0] == 2]
    for _, dj, _, b0, b1 in S_launches:
        for j in Ln["s_col"][b0:b1]:
            assert dims[j] == dj
            s = rhs[cs[j]:cs[j] + dj].copy()
            for p in range(A["fr_ptr"][j], A["fr_ptr"][j + 1]):
                dk, pk = Ln["fr_d"][p], Ln["fr_p"][p]
                k = int(np.searchsorted(plan.pstart, pk))
                assert plan.pstart[k] == pk and dims[k] == dk and y[k] is not None
                s -= F[A["fr_off"][p]:A["fr_off"][p] + dj * dk].reshape(dj, dk) @ y[k]
            y[j] = diag_solve(j, s, False)
    xs = [None] * N
    for _, dj, _, b0, b1 in reversed(S_launches):
        for j in Ln["s_col"][b0:b1]:
            s = y[j].copy()
            for p in range(A["bc_ptr"][j], A["bc_ptr"][j + 1]):
                di, pi = Ln["bc_d"][p], Ln["bc_p"][p]
                i = int(np.searchsorted(plan.pstart, pi))
                assert plan.pstart[i] == pi and dims[i] == di and xs[i] is not None
                s -= F[A["bc_off"][p]:A["bc_off"][p] + di * dj].reshape(di, dj).T @ xs[i]
            xs[j] = diag_solve(j, s, True)
    x = np.zeros_like(rhs)
    for j in range(N):
        x[cs[j]:cs[j] + dims[j]] = xs[j]
    return x


def random_block_spd(rng, sizes, fill):
    """Random block-sparse SPD matrix in the style of theseus/utils/sparse_matrix_utils.py:193-227 + AtA + I."""
    N = len(sizes)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    n = starts[-1]
    rows = max(3 * N, 8)
    J = np.zeros((rows * 2, n))
    for r in range(rows):
        vs = [v for v in range(N) if rng.random() < fill] or [int(rng.integers(N))]
        if len(vs) == 1 and N > 1:
            vs.append(int((vs[0] + 1 + rng.integers(N - 1)) % N))
        for v in set(vs):
            J[2 * r:2 * r + 2, starts[v]:starts[v + 1]] = rng.standard_normal((2, sizes[v]))
    M = J.T @ J + np.eye(n)
    nbr = [set([i]) for i in range(N)]
    for i in range(N):
        for j in range(N):
            if np.abs(M[starts[i]:starts[i + 1], starts[j]:starts[j + 1]]).max() > 0:
                nbr[i].add(j)
    ptrs, inds = [0], []
    for i in range(N):
        inds += sorted(nbr[i])
        ptrs.append(len(inds))
    return M, np.array(ptrs), np.array(inds)


@pytest.mark.parametrize("sizes,fill,ordering", [
    ([2, 3, 5, 2], 0.6, "mindeg"),             # the reference's literal test shape
    ([2, 3, 5, 2], 0.6, "natural"),
    ([6] * 12, 0.15, "mindeg"),
    ([1, 13, 2, 6, 3, 3, 6, 6, 4, 2, 5], 0.2, "mindeg"),   # param_size_range 1:13 (extlib/test_baspacho.py)
    ([3] * 30 + [6] * 5, 0.08, "mindeg"),      # bundle-adjustment-like: many small blocks tied to a few big ones
])
def test_plan_solves_system(sizes, fill, ordering):
    rng = np.random.default_rng(len(sizes) + int(fill * 100))
    M, ptrs, inds = random_block_spd(rng, sizes, fill)
    plan = analyze(np.array(sizes), ptrs, inds, ordering=ordering)
    assert sorted(plan.order.tolist()) == list(range(len(sizes)))
    rhs = rng.standard_normal(M.shape[0])
    x = run_plan_numpy(plan, M, rhs)
    assert np.abs(M @ x - rhs).max() < 1e-10   # the reference's criterion (extlib/test_baspacho.py:101-114)


@pytest.mark.parametrize("sizes,fill,ordering", [
    ([6] * 12, 0.15, "mindeg"),
    ([6] * 14, 0.9, "natural"),                  # dense: the last blocks have >= 8 update pairs (split-K launches)
    ([1, 2, 3, 6, 3, 3, 6, 6, 2, 1, 6], 0.25, "mindeg"),
    ([3] * 30 + [6] * 5, 0.08, "mindeg"),
])
def test_lane_work_lists_solve_system(sizes, fill, ordering):
    """The per-level / per-shape launch list of the batch-lane kernels, executed in order by a numpy interpreter."""
    rng = np.random.default_rng(len(sizes) + int(fill * 100) + 1)
    M, ptrs, inds = random_block_spd(rng, sizes, fill)
    plan = analyze(np.array(sizes), ptrs, inds, ordering=ordering)
    launches = plan.lane["launches"]
    assert launches.dtype == np.int32 and launches.shape[1] == 5
    if fill > 0.5:
        assert (launches[:, 0] == 3).any()
    rhs = rng.standard_normal(M.shape[0])
    x = run_lane_plan_numpy(plan, M, rhs)
    assert np.abs(M @ x - rhs).max() < 1e-10
    # every block of L is triangular-solved exactly once, every column substituted exactly once
    nT = sum(b1 - b0 for k, _, _, b0, b1 in launches if k == 1)
    nS = sum(b1 - b0 for k, _, _, b0, b1 in launches if k == 2)
    assert nT == len(plan.blk_off) and nS == plan.N


@pytest.mark.parametrize("sizes,fill,ordering,expect_tiles", [
    ([6] * 14, 0.9, "natural", True),      # one dense chain: the pieces after the first get their external updates from tiles
    ([6] * 40, 0.06, "mindeg", True),
    ([6] * 30 + [3] * 10, 0.07, "mindeg", True),   # mixed sizes: tiles that touch a 3-block stay on the per-block path
    ([3] * 30 + [6] * 5, 0.08, "mindeg", False),
])
def test_tiled_lane_lists_solve_system(sizes, fill, ordering, expect_tiles):
    """`lane_tiled` layout: flat tile arrays + reduced U lists (sparse.tile_lane_lists), executed by the numpy interpreter."""
    from theseus_b200.sparse import tile_lane_lists
    rng = np.random.default_rng(len(sizes) + int(fill * 100) + 7)
    M, ptrs, inds = random_block_spd(rng, sizes, fill)
    plan = analyze(np.array(sizes), ptrs, inds, ordering=ordering)
    lane, tiles = tile_lane_lists(plan)
    T = tiles["tile_tgt"].shape[0]
    assert tiles["tile_tgt"].shape == (T, 16) and tiles["step_ptr"].shape == (T + 1,) and tiles["step_src"].shape[1] == 8
    if expect_tiles:
        assert T > 0
    tiled_launches = [l for l in lane["launches"] if l[0] == 4]
    assert sum(l[4] - l[3] for l in tiled_launches) == T
    rhs = rng.standard_normal(M.shape[0])
    x = run_lane_plan_numpy(plan, M, rhs, lane=lane, tiles=tiles)
    assert np.abs(M @ x - rhs).max() < 1e-10
    # update pairs: every pair is done exactly once, by a tile step or by a U item
    n_u = int((lane["u_p1"] - lane["u_p0"]).sum())
    n_t = 0
    for t in range(T):
        tg = tiles["tile_tgt"][t]
        for st in range(tiles["step_ptr"][t], tiles["step_ptr"][t + 1]):
            src = tiles["step_src"][st]
            n_t += sum(1 for w in range(16) if tg[w] >= 0 and src[w // 4] >= 0 and src[4 + w % 4] >= 0)
    assert n_u + n_t == len(plan.arrays["up_a"])
    print(f"tiles={T} tile updates={n_t} per-block updates={n_u}")


@pytest.mark.parametrize("sizes,fill,ordering", [
    ([2, 3, 5, 2], 0.6, "mindeg"), ([6] * 14, 0.9, "natural"), ([1, 2, 3, 6, 3, 3, 6, 6, 2, 1, 6], 0.25, "mindeg"),
    ([3] * 30 + [6] * 5, 0.08, "mindeg"), ([6] * 120, 0.03, "mindeg"),
])
def test_native_symbolic_equals_python_specification(sizes, fill, ordering):
    """thb_symbolic_create (C++, csrc/thb_symbolic.cu) against sparse.analyze_py: every array of both numeric back ends, the
    ordering, the structure, the statistics -- bit for bit (integer data)."""
    rng = np.random.default_rng(len(sizes))
    M, ptrs, inds = random_block_spd(rng, sizes, fill)
    a, b = analyze(np.array(sizes), ptrs, inds, ordering=ordering), analyze_py(np.array(sizes), ptrs, inds, ordering=ordering)
    assert set(a.arrays) == set(b.arrays) and set(a.lane) == set(b.lane)
    for k in a.arrays:
        assert a.arrays[k].dtype == b.arrays[k].dtype and np.array_equal(a.arrays[k], b.arrays[k]), k
    for k in a.lane:
        assert a.lane[k].dtype == b.lane[k].dtype and a.lane[k].shape == b.lane[k].shape and np.array_equal(a.lane[k], b.lane[k]), k
    for k in ("order", "pos", "dims", "col_start", "pstart", "level", "blk_off", "blk_rows", "blk_cols", "winv_off"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert a.blk_index == b.blk_index and a.stats == b.stats
    assert np.array_equal(a.chain_of, b.chain_of) and np.array_equal(a.chain_level, b.chain_level)
    assert all(np.array_equal(x, y) for x, y in zip(a.struct, b.struct))
    assert (a.N, a.n, a.data_size, a.winv_size) == (b.N, b.n, b.data_size, b.winv_size)


def run_root_split_numpy(plan, sp, M, rhs, lane=None, tiles=None):
    """Interpret sparse.root_split: lane work lists for the columns below the cut, then the dense root (Schur complement assembled from
    the bottom columns' update pairs, numpy Cholesky standing in for the dense DMMA kernel), substitutions split the same way."""
    A, Ln = plan.arrays, (sp["bottom"] if lane is None else lane)   # `lane` (tile_lane_lists with a split): root assembly included
    N, dims, cs, cut = plan.N, plan.dims, plan.col_start, sp["cut"]
    F = np.zeros(plan.data_size)
    for (i, j), t in plan.blk_index.items():
        blk = M[cs[i]:cs[i] + dims[i], cs[j]:cs[j] + dims[j]]
        F[plan.blk_off[t]:plan.blk_off[t] + blk.size] = blk.reshape(-1)
    DL = np.zeros(plan.winv_size)

    def update(tgt, di, dj, p0, p1):
        T = F[tgt:tgt + di * dj].reshape(di, dj).copy()
        for p in range(p0, p1):
            dk = A["up_k"][p]
            T -= F[A["up_a"][p]:A["up_a"][p] + di * dk].reshape(di, dk) @ F[A["up_b"][p]:A["up_b"][p] + dj * dk].reshape(dj, dk).T
        F[tgt:tgt + di * dj] = T.reshape(-1)

    for kind, di, dj, b0, b1 in Ln["launches"]:
        if kind in (0, 3):
            for e in range(b0, b1):
                update(Ln["u_tgt"][e], di, dj, Ln["u_p0"][e], Ln["u_p1"][e])
        elif kind == 4:
            apply_tiles_numpy(F, tiles, b0, b1, di, dj)
        elif kind == 1:
            writes = {}
            for e in range(b0, b1):
                off, dg, dl = Ln["t_off"][e], Ln["t_diag"][e], Ln["t_dl"][e]
                D = np.tril(F[dg:dg + dj * dj].reshape(dj, dj))
                L = np.linalg.cholesky(D + np.tril(D, -1).T)
                if off == dg:
                    Lr = L.copy(); Lr[np.arange(dj), np.arange(dj)] = 1.0 / np.diag(L); writes[dl] = Lr
                else:
                    F[off:off + di * dj] = np.linalg.solve(L, F[off:off + di * dj].reshape(di, dj).T).T.reshape(-1)
            for dl, Lr in writes.items():
                DL[dl:dl + Lr.size] = Lr.reshape(-1)
    # root: assemble, copy to a dense matrix, factor densely
    blk_dims = {int(plan.blk_off[t]): (int(dims[i]), int(dims[j])) for (i, j), t in plan.blk_index.items()}
    for e in range(len(sp["ru_tgt"]) if lane is None else 0):
        di, dj = blk_dims[int(sp["ru_tgt"][e])]
        update(sp["ru_tgt"][e], di, dj, sp["ru_p0"][e], sp["ru_p1"][e])
    nt = sp["root_dof"]
    S = np.zeros((nt, nt))
    for e in range(len(sp["rb_off"])):
        di, dj, r0, c0, off = sp["rb_di"][e], sp["rb_dj"][e], sp["rb_row"][e], sp["rb_col"][e], sp["rb_off"][e]
        S[r0:r0 + di, c0:c0 + dj] = F[off:off + di * dj].reshape(di, dj)
    S = np.tril(S) + np.tril(S, -1).T
    Lroot = np.linalg.cholesky(S)

    def diag_solve(j, s, transpose):
        d = dims[j]
        Lr = DL[A["winv_off"][j]:A["winv_off"][j] + d * d].reshape(d, d).copy()
        Lr[np.arange(d), np.arange(d)] = 1.0 / np.diag(Lr)
        return np.linalg.solve(Lr.T if transpose else Lr, s)

    y = [None] * N
    S_l = [l for l in Ln["launches"] if l[0] == 2]
    for _, dj, _, b0, b1 in S_l:
        for j in Ln["s_col"][b0:b1]:
            assert j < cut
            s = rhs[cs[j]:cs[j] + dj].copy()
            for p in range(A["fr_ptr"][j], A["fr_ptr"][j + 1]):
                k = A["fr_k"][p]
                s -= F[A["fr_off"][p]:A["fr_off"][p] + dj * dims[k]].reshape(dj, dims[k]) @ y[k]
            y[j] = diag_solve(j, s, False)
    sr = np.zeros(nt)
    for q, j in enumerate(range(cut, N)):
        s = rhs[cs[j]:cs[j] + dims[j]].copy()
        for p in range(sp["rf_p0"][q], sp["rf_p1"][q]):
            k = A["fr_k"][p]
            assert k < cut
            s -= F[A["fr_off"][p]:A["fr_off"][p] + dims[j] * dims[k]].reshape(dims[j], dims[k]) @ y[k]
        sr[plan.pstart[j] - sp["root_start"]:plan.pstart[j] - sp["root_start"] + dims[j]] = s
    xr = np.linalg.solve(Lroot.T, np.linalg.solve(Lroot, sr))
    xs = [None] * N
    for j in range(cut, N):
        xs[j] = xr[plan.pstart[j] - sp["root_start"]:plan.pstart[j] - sp["root_start"] + dims[j]]
    for _, dj, _, b0, b1 in reversed(S_l):
        for j in Ln["s_col"][b0:b1]:
            s = y[j].copy()
            for p in range(A["bc_ptr"][j], A["bc_ptr"][j + 1]):
                i = A["bc_i"][p]
                s -= F[A["bc_off"][p]:A["bc_off"][p] + dims[i] * dj].reshape(dims[i], dj).T @ xs[i]
            xs[j] = diag_solve(j, s, True)
    x = np.zeros_like(rhs)
    for j in range(N):
        x[cs[j]:cs[j] + dims[j]] = xs[j]
    return x


def test_root_split_solves_system():
    """sparse.root_split (dense trailing block + lane lists below it) on a pose-graph-like structure with a long top chain."""
    rng = np.random.default_rng(21)
    N = 40
    sizes = [6] * N
    # ring + chords: minimum degree ends with a dense separator chain
    starts = np.arange(N + 1) * 6
    rows = []
    for i in range(N):
        for jn in ((i + 1) % N, (i + 7) % N):
            J = np.zeros((3, 6 * N)); J[:, starts[i]:starts[i + 1]] = rng.standard_normal((3, 6)); J[:, starts[jn]:starts[jn + 1]] = rng.standard_normal((3, 6))
            rows.append(J)
    Jm = np.concatenate(rows, 0)
    M = Jm.T @ Jm + np.eye(6 * N)
    nbr = [set([i]) for i in range(N)]
    for i in range(N):
        for jn in ((i + 1) % N, (i + 7) % N):
            nbr[i].add(jn); nbr[jn].add(i)
    ptrs, inds = [0], []
    for i in range(N):
        inds += sorted(nbr[i]); ptrs.append(len(inds))
    plan = analyze(np.array(sizes), np.array(ptrs), np.array(inds))
    sp = root_split(plan, max_root_dof=1024, min_root_cols=4)
    assert sp is not None and 4 <= N - sp["cut"] and sp["root_dof"] == 6 * (N - sp["cut"])
    assert len(sp["ru_tgt"]) > 0 and (sp["pair_k"][sp["ru_p0"][0]:sp["ru_p1"][0]] < sp["cut"]).all()
    rhs = rng.standard_normal(6 * N)
    x = run_root_split_numpy(plan, sp, M, rhs)
    assert np.abs(M @ x - rhs).max() < 1e-9
    # smaller root on request; no root when the tail is too short
    sp2 = root_split(plan, max_root_dof=24, min_root_cols=4)
    assert sp2 is not None and sp2["root_dof"] == 24
    assert np.abs(M @ run_root_split_numpy(plan, sp2, M, rhs) - rhs).max() < 1e-9
    assert root_split(plan, max_root_dof=12, min_root_cols=4) is None
    # tiled bottom + tiled root assembly (layout lane_tiled_root): same answer from the flat lists
    from theseus_b200.sparse import tile_lane_lists
    for s_ in (sp, sp2):
        lane, tiles = tile_lane_lists(plan, s_)
        tu = [l for l in lane["launches"] if l[0] == 4]
        assert tu and tuple(lane["launches"][-1][:1]) in ((4,), (0,), (3,))   # the assembly closes the list
        assert all(j < s_["cut"] for l in lane["launches"] if l[0] == 2 for j in lane["s_col"][l[3]:l[4]])
        assert np.abs(M @ run_root_split_numpy(plan, s_, M, rhs, lane=lane, tiles=tiles) - rhs).max() < 1e-9
        # pairs: bottom columns complete, root blocks only their k < cut part
        n_t = 0
        for t in range(tiles["tile_tgt"].shape[0]):
            tg = tiles["tile_tgt"][t]
            for st in range(tiles["step_ptr"][t], tiles["step_ptr"][t + 1]):
                src = tiles["step_src"][st]
                n_t += sum(1 for w in range(16) if tg[w] >= 0 and src[w // 4] >= 0 and src[4 + w % 4] >= 0)
        n_all = int((s_["bottom"]["u_p1"] - s_["bottom"]["u_p0"]).sum() + (s_["ru_p1"] - s_["ru_p0"]).sum())
        assert n_t + int((lane["u_p1"] - lane["u_p0"]).sum()) == n_all


def run_chain_tiles_numpy(plan, ct, M, rhs):
    """Interpret sparse.chain_tiles: per level (1) the external updates of the tiles due at this level, each source block of a step
    fetched ONCE per tile, (2) the internal updates of the level's columns, (3) the diagonal / triangular stage; then the ordinary
    substitutions (run_plan_numpy's lists)."""
    A = plan.arrays
    N, dims, cs = plan.N, plan.dims, plan.col_start
    F = np.zeros(plan.data_size)
    for (i, j), t in plan.blk_index.items():
        blk = M[cs[i]:cs[i] + dims[i], cs[j]:cs[j] + dims[j]]
        F[plan.blk_off[t]:plan.blk_off[t] + blk.size] = blk.reshape(-1)
    blk_ij = {t: ij for ij, t in plan.blk_index.items()}

    def B(t):
        i, j = blk_ij[t]
        return F[plan.blk_off[t]:plan.blk_off[t] + dims[i] * dims[j]].reshape(dims[i], dims[j])

    def setB(t, V):
        F[plan.blk_off[t]:plan.blk_off[t] + V.size] = V.reshape(-1)

    W = np.zeros(plan.winv_size)
    done_pairs = 0
    for lv, L in enumerate(ct["per_level"]):
        for tile in L["tiles"]:
            acc = {t: B(t).copy() for (_, _, t) in tile["targets"]}
            for (k, ro, co) in tile["steps"]:
                staged = {x: B(x).copy() for x in set(ro + co) if x >= 0}     # one fetch per source block and step
                for (a, b, t) in tile["targets"]:
                    if ro[a] >= 0 and co[b] >= 0:
                        acc[t] -= staged[ro[a]] @ staged[co[b]].T
                        done_pairs += 1
            for t, V in acc.items():
                setB(t, V)
        for (t, p0, p1) in L["u_int"]:
            V = B(t).copy()
            i, j = blk_ij[t]
            for p in range(p0, p1):
                dk = A["up_k"][p]
                V -= F[A["up_a"][p]:A["up_a"][p] + dims[i] * dk].reshape(dims[i], dk) @ F[A["up_b"][p]:A["up_b"][p] + dims[j] * dk].reshape(dims[j], dk).T
                done_pairs += 1
            setB(t, V)
        for j in L["cols"]:
            tj = plan.blk_index[(j, j)]
            D = np.tril(B(tj))
            Lj = np.linalg.cholesky(D + np.tril(D, -1).T)
            setB(tj, Lj)
            W[A["winv_off"][j]:A["winv_off"][j] + dims[j] ** 2] = np.linalg.inv(Lj).reshape(-1)
            for i in plan.struct[j]:
                t = plan.blk_index[(int(i), j)]
                setB(t, np.linalg.solve(Lj, B(t).T).T)
    assert done_pairs == int(plan.stats["num_updates"]) == ct["updates"] + sum(p1 - p0 for L in ct["per_level"] for (_, p0, p1) in L["u_int"])
    # substitutions exactly as run_plan_numpy (item-layout lists; W = inverse diagonal factors)
    nlev = len(A["s_ptr"]) - 1
    y = [rhs[cs[j]:cs[j] + dims[j]].copy() for j in range(N)]
    for lv in range(nlev):
        for e in range(A["s_ptr"][lv], A["s_ptr"][lv + 1]):
            j = A["s_col"][e]
            s = y[j].copy()
            for p in range(A["fr_ptr"][j], A["fr_ptr"][j + 1]):
                k = A["fr_k"][p]
                s -= F[A["fr_off"][p]:A["fr_off"][p] + dims[j] * dims[k]].reshape(dims[j], dims[k]) @ y[k]
            y[j] = W[A["winv_off"][j]:A["winv_off"][j] + dims[j] ** 2].reshape(dims[j], dims[j]) @ s
    for lv in reversed(range(nlev)):
        for e in range(A["s_ptr"][lv], A["s_ptr"][lv + 1]):
            j = A["s_col"][e]
            s = y[j].copy()
            for p in range(A["bc_ptr"][j], A["bc_ptr"][j + 1]):
                i = A["bc_i"][p]
                s -= F[A["bc_off"][p]:A["bc_off"][p] + dims[i] * dims[j]].reshape(dims[i], dims[j]).T @ y[i]
            y[j] = W[A["winv_off"][j]:A["winv_off"][j] + dims[j] ** 2].reshape(dims[j], dims[j]).T @ s
    x = np.zeros_like(rhs)
    for j in range(N):
        x[cs[j]:cs[j] + dims[j]] = y[j]
    return x


@pytest.mark.parametrize("sizes,fill,width,rows", [([6] * 30, 0.08, 4, 4), ([3, 6] * 14, 0.12, 2, 3), ([6] * 16, 0.9, 4, 4), ([2, 3, 6, 1] * 8, 0.1, 8, 2)])
def test_chain_tiles_schedule_solves_system(sizes, fill, width, rows):
    """The tiled-update schedule (sparse.chain_tiles: pieces of chains, external pairs per (piece, row tile) with every source block
    fetched once per step, internal pairs per column) performs exactly the plan's update pairs and factorises correctly."""
    rng = np.random.default_rng(len(sizes) * 7 + width)
    M, ptrs, inds = random_block_spd(rng, sizes, fill)
    plan = analyze(np.array(sizes), ptrs, inds)
    ct = chain_tiles(plan, max_width=width, tile_rows=rows)
    pf = ct["piece_first"]
    for j in range(plan.N):   # pieces are runs of <= width columns inside one chain
        assert pf[j] <= j < pf[j] + width and plan.chain_of[pf[j]] == plan.chain_of[j]
    rhs = rng.standard_normal(M.shape[0])
    x = run_chain_tiles_numpy(plan, ct, M, rhs)
    assert np.abs(M @ x - rhs).max() < 1e-9
    if ct["updates"]:
        assert ct["block_loads"] <= 2 * ct["updates"]     # never worse than streaming both operands of every update


def test_chains_are_fundamental_supernodes():
    """Chain partition (the unit of the round-2 numeric phase): consecutive columns, nested structures, a tree of chains."""
    rng = np.random.default_rng(9)
    M, ptrs, inds = random_block_spd(rng, [6] * 60, 0.05)
    p = analyze(np.array([6] * 60), ptrs, inds)
    N = p.N
    parent = [int(p.struct[j][0]) if len(p.struct[j]) else -1 for j in range(N)]
    assert p.chain_of[0] == 0 and (np.diff(p.chain_of) >= 0).all() and (np.diff(p.chain_of) <= 1).all()   # chains are runs of columns
    for j in range(N - 1):
        if p.chain_of[j] == p.chain_of[j + 1]:
            assert parent[j] == j + 1 and set(p.struct[j].tolist()) == {j + 1} | set(p.struct[j + 1].tolist())
            assert sum(1 for q in range(N) if parent[q] == j + 1) == 1
    for j in range(N):   # a chain sits above all chains that feed it
        if parent[j] >= 0 and p.chain_of[parent[j]] != p.chain_of[j]:
            assert p.chain_level[p.chain_of[parent[j]]] > p.chain_level[p.chain_of[j]]
    assert p.stats["num_chains"] == p.chain_of.max() + 1 <= N and p.stats["chain_levels"] <= p.stats["levels"]


def test_minimum_degree_eliminates_leaves_first_and_is_deterministic():
    # star + chain: the leaves of the star must go before the hub
    N = 8
    nbr = {0: [1, 2, 3, 4, 5], 5: [6], 6: [7]}
    adj = [set([i]) for i in range(N)]
    for a, bs in nbr.items():
        for b in bs:
            adj[a].add(b); adj[b].add(a)
    ptrs, inds = [0], []
    for i in range(N):
        inds += sorted(adj[i]); ptrs.append(len(inds))
    o1 = minimum_degree_order(N, np.array(ptrs), np.array(inds), np.ones(N, dtype=np.int64))
    o2 = minimum_degree_order(N, np.array(ptrs), np.array(inds), np.ones(N, dtype=np.int64))
    assert o1.tolist() == o2.tolist()
    assert list(o1).index(0) > list(o1).index(1)


def test_fill_reducing_quality_on_pose_graph_ring():
    # 60-node ring with chords: minimum degree must beat the natural order clearly (ordering is worth 3-5x)
    N = 60
    adj = [set([i, (i + 1) % N, (i - 1) % N]) for i in range(N)]
    for i in range(0, N, 7):
        j = (i * 3 + 11) % N
        adj[i].add(j); adj[j].add(i)
    ptrs, inds = [0], []
    for i in range(N):
        inds += sorted(adj[i]); ptrs.append(len(inds))
    sizes = np.full(N, 6)
    a = analyze(sizes, np.array(ptrs), np.array(inds), "mindeg")
    b = analyze(sizes, np.array(ptrs), np.array(inds), "natural")
    assert a.stats["nnz_L"] < 0.8 * b.stats["nnz_L"]


def run_piece_solve_numpy(plan, ps, F, DL, rhs):
    """Interpret sparse.piece_solve_lists on a finished factor (F = blocks of L below the diagonal, DL = diagonal factors with reciprocal
    diagonal, as the lane kernels leave them): forward then backward, piece launches in order; inside a launch the pieces are independent
    (checked: a piece only reads y / x entries written by EARLIER launches or by itself)."""
    A = plan.arrays
    N, dims, cs = plan.N, plan.dims, plan.col_start
    cut = ps["cut"]

    def diag(j):
        d = dims[j]
        Lr = DL[A["winv_off"][j]:A["winv_off"][j] + d * d].reshape(d, d).copy()
        Lr[np.arange(d), np.arange(d)] = 1.0 / np.diag(Lr)
        return np.tril(Lr)

    y = [None] * N
    for (lv, d, b0, b1) in ps["launches"]:
        written = []
        for p in ps["order"][b0:b1]:
            j0, w = ps["first"][p], ps["width"][p]
            s = []
            for jj in range(j0, j0 + w):                          # external parts: independent of each other
                assert dims[jj] == d
                acc = rhs[cs[jj]:cs[jj] + d].copy()
                for q in range(A["fr_ptr"][jj], ps["fr_ext_end"][jj]):
                    k = A["fr_k"][q]
                    assert y[k] is not None and k not in written
                    acc -= F[A["fr_off"][q]:A["fr_off"][q] + d * dims[k]].reshape(d, dims[k]) @ y[k]
                s.append(acc)
            for c, jj in enumerate(range(j0, j0 + w)):            # internal triangle, in order
                for q in range(ps["fr_ext_end"][jj], A["fr_ptr"][jj + 1]):
                    k = A["fr_k"][q]
                    assert j0 <= k < jj
                    s[c] -= F[A["fr_off"][q]:A["fr_off"][q] + d * d].reshape(d, d) @ y[k]
                y[jj] = np.linalg.solve(diag(jj), s[c])
            written.extend(range(j0, j0 + w))
    return y


def run_piece_backward_numpy(plan, ps, F, DL, y, x_known):
    A = plan.arrays
    N, dims = plan.N, plan.dims

    def diag(j):
        d = dims[j]
        Lr = DL[A["winv_off"][j]:A["winv_off"][j] + d * d].reshape(d, d).copy()
        Lr[np.arange(d), np.arange(d)] = 1.0 / np.diag(Lr)
        return np.tril(Lr)

    x = list(x_known)
    for (lv, d, b0, b1) in ps["launches"][::-1]:
        written = []
        for p in ps["order"][b0:b1]:
            j0, w = ps["first"][p], ps["width"][p]
            j1 = j0 + w - 1
            s = [y[jj].copy() for jj in range(j0, j1 + 1)]
            n_ext = A["bc_ptr"][j1 + 1] - ps["bc_int_end"][j1]
            for e in range(n_ext):                                 # every external x_i is fetched once for the whole piece
                i = A["bc_i"][ps["bc_int_end"][j1] + e]
                assert x[i] is not None and i not in written
                for c, jj in enumerate(range(j0, j1 + 1)):
                    q = ps["bc_int_end"][jj] + e
                    assert A["bc_i"][q] == i
                    s[c] -= F[A["bc_off"][q]:A["bc_off"][q] + dims[i] * d].reshape(dims[i], d).T @ x[i]
            for c, jj in reversed(list(enumerate(range(j0, j1 + 1)))):
                for q in range(A["bc_ptr"][jj], ps["bc_int_end"][jj]):
                    i = A["bc_i"][q]
                    assert jj < i <= j1
                    s[c] -= F[A["bc_off"][q]:A["bc_off"][q] + d * d].reshape(d, d).T @ x[i]
                x[jj] = np.linalg.solve(diag(jj).T, s[c])
            written.extend(range(j0, j1 + 1))
    return x


def _factor_numpy(plan, M):
    """Block Cholesky in the plan's storage, as the lane kernels leave it."""
    N, dims, cs = plan.N, plan.dims, plan.col_start
    order_start = plan.pstart
    n = int(dims.sum())
    perm = np.concatenate([np.arange(cs[j], cs[j] + dims[j]) for j in range(N)])
    L = np.linalg.cholesky(M[np.ix_(perm, perm)])
    F = np.zeros(plan.data_size)
    DL = np.zeros(plan.winv_size)
    for (i, j), t in plan.blk_index.items():
        blk = L[order_start[i]:order_start[i] + dims[i], order_start[j]:order_start[j] + dims[j]]
        if i == j:
            Lr = blk.copy()
            Lr[np.arange(dims[j]), np.arange(dims[j])] = 1.0 / np.diag(blk)
            DL[plan.arrays["winv_off"][j]:plan.arrays["winv_off"][j] + blk.size] = Lr.reshape(-1)
        else:
            F[plan.blk_off[t]:plan.blk_off[t] + blk.size] = blk.reshape(-1)
    return F, DL, perm, L


@pytest.mark.parametrize("sizes,fill,ordering,width", [
    ([6] * 14, 0.9, "natural", 4), ([6] * 40, 0.06, "mindeg", 4), ([6] * 40, 0.06, "mindeg", 2), ([6] * 30 + [3] * 10, 0.07, "mindeg", 4),
    ([1, 2, 3, 6, 3, 3, 6, 6, 2, 1, 6], 0.25, "mindeg", 8), ([3] * 30 + [6] * 5, 0.08, "mindeg", 4)])
def test_piece_solve_schedule_solves_system(sizes, fill, ordering, width):
    """sparse.piece_solve_lists: supernodal (chain-piece) forward / backward substitution schedule against a dense solve."""
    from theseus_b200.sparse import piece_solve_lists
    rng = np.random.default_rng(len(sizes) + int(fill * 100) + 11 * width)
    M, ptrs, inds = random_block_spd(rng, sizes, fill)
    plan = analyze(np.array(sizes), ptrs, inds, ordering=ordering)
    ps = piece_solve_lists(plan, max_width=width)
    P = len(ps["first"])
    assert int(ps["width"].sum()) == plan.N and ps["width"].max() <= width
    assert sum(b1 - b0 for (_, _, b0, b1) in ps["launches"]) == P and sorted(ps["order"].tolist()) == list(range(P))
    nlev = int(ps["level"].max()) + 1
    assert nlev <= int(plan.level.max()) + 1
    F, DL, perm, L = _factor_numpy(plan, M)
    rhs = rng.standard_normal(M.shape[0])
    y = run_piece_solve_numpy(plan, ps, F, DL, rhs)
    x = run_piece_backward_numpy(plan, ps, F, DL, y, [None] * plan.N)
    xs = np.zeros_like(rhs)
    for j in range(plan.N):
        xs[plan.col_start[j]:plan.col_start[j] + plan.dims[j]] = x[j]
    assert np.abs(M @ xs - rhs).max() < 1e-9
    print(f"columns {plan.N}, pieces {P}, piece levels {nlev} (column levels {int(plan.level.max()) + 1})")
