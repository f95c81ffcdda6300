"""Symbolic analysis of the batched MULTIFRONTAL (supernodal) block-sparse Cholesky -- layout "front" of BaspachoSparseSolver.

Replaces what the reference delegates to BaSpaCho::createSolver behind SymbolicDecomposition(param_size, block_ptrs, block_inds, dev)
(theseus/extlib/baspacho_solver.cpp:259-319: fill-reducing ordering, supernodes, factor layout); same inputs, internals unobservable in
the reference ("parity unpinned") -- the un-permuted solution is what is checked.

Steps (host, batch independent, once per structure):
  1. ordering: nested dissection of the block graph (level-set separators from pseudo-peripheral nodes) and minimum degree
     (sparse.minimum_degree_order); the one with fewer factorisation flops wins (pose graphs: ND; bundle adjustment: min degree);
  2. elimination tree + column structures;
  3. supernodes -> FRONTS by relaxed amalgamation (children merged into the parent while the dense flops grow little);
  4. per front s: w_s pivot scalars (contiguous in the permuted vector), b_s border rows, r_s = w_s + b_s; the factor panel
     L[rows_s, cols_s] is a dense row-major r_s x w_s matrix at panel_off[s] of one item's factor storage;
  5. depth schedule: all children of a front sit exactly one depth below it, so update matrices (and the forward substitution's
     border vectors) live for one step in a ping-pong arena indexed by depth parity;
  6. child -> parent relative row maps, per-class launch lists (small fronts: one CTA per (front, item) in shared memory;
     big fronts: assembled in global memory and factored by the DMMA dense kernel in partial mode).
`execute_numpy` runs the very arrays the kernels consume (tests/test_frontal.py).
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

SMALL_CLASSES = (48, 96, 1 << 30)  # front sizes (scalar rows) of the shared-memory kernel's thread-count classes: 64 / 128 / 256 threads
BIG_TW, BIG_TM = 64, 128          # block-column width / row-tile height of the dense DMMA kernel (thb_chol_dense.cu)


# ------------------------------------------------------------------------------------------------ ordering
def _bfs_levels(adj, start, mask, stamp, mark):
    mark[start] = True
    levels, seen = [[start]], [start]
    while True:
        nxt = []
        for u in levels[-1]:
            for v in adj[u]:
                if mask[v] == stamp and not mark[v]:
                    mark[v] = True
                    nxt.append(v)
                    seen.append(v)
        if not nxt:
            break
        levels.append(nxt)
    for v in seen:
        mark[v] = False
    return levels, seen


def nested_dissection_order(N: int, ptrs: np.ndarray, inds: np.ndarray, leaf: int = 8) -> np.ndarray:
    """order[k] = variable eliminated k-th.  Recursive bisection by the middle level set of a BFS from a pseudo-peripheral node
    (George's automatic nested dissection), separator nodes that do not touch the far side moved back, separators ordered last;
    deterministic."""
    adj = [[int(x) for x in inds[ptrs[i]:ptrs[i + 1]] if int(x) != i] for i in range(N)]
    mask = np.zeros(N, dtype=np.int64)
    mark = np.zeros(N, dtype=bool)
    order: List[int] = []
    stamp = [0]
    # explicit stack, emitting separators AFTER both halves: entries are ("nodes", list) or ("emit", list)
    work = [("nodes", list(range(N)))]
    while work:
        kind, nodes = work.pop()
        if kind == "emit":
            order.extend(nodes)
            continue
        if not nodes:
            continue
        stamp[0] += 1
        s = stamp[0]
        for v in nodes:
            mask[v] = s
        if len(nodes) <= leaf:
            order.extend(sorted(nodes, key=lambda v: (sum(1 for x in adj[v] if mask[x] == s), v)))
            for v in nodes:
                mask[v] = 0
            continue
        remaining = set(nodes)
        comps = []
        for v in nodes:
            if v in remaining:
                _, seen = _bfs_levels(adj, v, mask, s, mark)
                comps.append(seen)
                remaining.difference_update(seen)
        if len(comps) > 1:
            for c in reversed(comps):
                work.append(("nodes", c))
            continue
        levels, _ = _bfs_levels(adj, nodes[0], mask, s, mark)
        for _ in range(4):
            cand = min(levels[-1], key=lambda v: (sum(1 for x in adj[v] if mask[x] == s), v))
            l2, _ = _bfs_levels(adj, cand, mask, s, mark)
            if len(l2) > len(levels):
                levels = l2
            else:
                break
        if len(levels) < 3:
            order.extend(sorted(nodes))
            for v in nodes:
                mask[v] = 0
            continue
        tot = len(nodes)
        cum = np.cumsum([len(l) for l in levels])
        best, bi = None, None
        for i in range(1, len(levels) - 1):
            left, right = int(cum[i - 1]), tot - int(cum[i])
            bal = min(left, right) / max(left, right, 1)
            if bal < 0.4:
                continue
            score = (len(levels[i]), -bal)
            if best is None or score < best:
                best, bi = score, i
        if bi is None:
            bi = min(max(int(np.argmin(np.abs(cum - tot / 2))), 1), len(levels) - 2)
        right = [v for l in levels[bi + 1:] for v in l]
        left = [v for l in levels[:bi] for v in l]
        rs = set(right)
        sep = []
        for v in levels[bi]:
            if any((x in rs) for x in adj[v]):
                sep.append(v)
            else:
                left.append(v)
        for v in sep:
            mask[v] = 0
        work.append(("emit", sorted(sep)))
        work.append(("nodes", right))
        work.append(("nodes", left))
    assert len(order) == N
    return np.array(order, dtype=np.int64)


def _column_structures(N, ptrs, inds, order):
    """Elimination tree parent[] and the sorted below-diagonal block structure of every column, in elimination positions."""
    pos = np.empty(N, dtype=np.int64)
    pos[order] = np.arange(N)
    parent = -np.ones(N, dtype=np.int64)
    struct: List[Optional[set]] = [None] * N
    children: List[List[int]] = [[] for _ in range(N)]
    for j in range(N):
        v = int(order[j])
        s = set(int(pos[x]) for x in inds[ptrs[v]:ptrs[v + 1]] if pos[x] > j)
        for c in children[j]:
            s.update(struct[c])
        s.discard(j)
        struct[j] = s
        if s:
            p = min(s)
            parent[j] = p
            children[p].append(j)
    return pos, parent, [np.array(sorted(s), dtype=np.int64) for s in struct]


def _flops_of(struct, dims):
    """Factorisation flops (mul + add) of the block column structures."""
    fl = 0.0
    for j, s in enumerate(struct):
        r = float(sum(dims[i] for i in s))
        d = float(dims[j])
        fl += d ** 3 / 3.0 + d * d * r + d * r * r
    return fl


STRIPE = 32                          # rows of the update matrix produced at a time by the shared-memory kernel (thb_front.cu FRONT_STRIPE)
SMALL_SMEM_LIMIT = 220 * 1024
SMEM_BUCKETS = (27 * 1024, 36 * 1024, 55 * 1024, 74 * 1024, 112 * 1024)   # 8, 6, 4, 3, 2 CTAs per SM (227 KB usable), then 1


def _pad_ld(x: int) -> int:
    return ((x + 11) // 16) * 16 + 4


def small_smem_bytes(w: int, b: int, nchildren: int = 0) -> int:
    """Dynamic shared memory of front_small_kernel for a front (thb_front.cu: front_smem_doubles): the padded panel, the inverse of one
    8 x 8 diagonal block and the children's int32 inverse maps.  (The update matrix is never resident: its tiles go from registers to
    global memory.)"""
    b16, w8 = (b + 15) & ~15, (w + 7) & ~7
    return ((w8 + b16 + 8) * _pad_ld(w8) + 8 * 20 + 3 * SMALL_MAX_CHILDREN + 2 + (min(nchildren, SMALL_MAX_CHILDREN) * (w + b) + 1) // 2) * 8


SPLIT_MAX_W = int(__import__("os").environ.get("THB_FRONT_SPLIT_W", "96"))   # pivot columns of one piece when a wide supernode is split into a chain
SPLIT_ROOT_W = 256       # borderless fronts wider than this stay whole (dense path)
SMALL_MAX_CHILDREN = 8   # thb_front.cu FRONT_MAX_CHILDREN: the gather kernel keeps its children's descriptors in registers
SMALL_MAX_W = 192    # pivot block columns of 8 are factored one after the other inside the CTA: wider pivot blocks go to the dense kernel


def _front_cost(w, b):
    return w ** 3 / 3.0 + w * w * b + w * b * b


# ------------------------------------------------------------------------------------------------ the plan
@dataclass
class FrontPlan:
    N: int
    n: int
    param_size: np.ndarray
    order: np.ndarray          # [N] order[k] = original variable at elimination position k
    pos: np.ndarray            # [N]
    dims: np.ndarray           # [N] per position
    pstart: np.ndarray         # [N] scalar start in the permuted vector, per position
    col_start: np.ndarray      # [N] scalar start in the ORIGINAL column layout, per position
    perm: np.ndarray           # [n] perm[p] = original scalar column of permuted scalar p
    S: int
    arrays: Dict[str, np.ndarray]   # flat per-front arrays consumed by the kernels (see build_front_plan)
    launches: np.ndarray       # [num_launches, 12] int64, thb200.h THB_FRONT_LAUNCH_COLS; depth descending
    data_size: int             # doubles of one item's factor storage (sum of panels)
    arena_size: int            # doubles of one item's update-matrix arena (one parity)
    varena_size: int           # doubles of one item's border-vector arena (one parity)
    stats: Dict[str, float] = field(default_factory=dict)
    front_of_pos: np.ndarray = None
    border_rows: List[np.ndarray] = None   # per front: permuted scalar indices of the border rows

    def gram_compact_offsets(self, blocks_dims):
        """Compact AtA block storage for the panel-map flow: `blocks_dims` = [(a, b, di, dj)] in the Gram plan's block order (original
        variables, pos[a] >= pos[b]).  Fills arrays["pmap"] (panel offset -> compact offset, -1 = fill-in) and returns
        (callable for structure.build_gram_plan, doubles per item)."""
        f = self.gram_out_offsets()
        pmap = self.arrays["pmap"]
        pmap[:] = -1
        table, off = {}, 0
        for (a, b, di, dj) in blocks_dims:
            po, ld, _ = f(a, b)
            idx = po + np.arange(di)[:, None] * ld + np.arange(dj)[None, :]
            pmap[idx] = off + np.arange(di)[:, None] * dj + np.arange(dj)[None, :]
            table[(a, b)] = (off, dj, -1)
            off += di * dj
        return (lambda a, b: table[(int(a), int(b))]), off

    def gram_out_offsets(self):
        """Callable for structure.build_gram_plan: where block (a, b) of AtA (original variables, pos[a] >= pos[b]) lands in the
        panel storage: (offset, leading dimension, mirror=-1)."""
        A = self.arrays
        first, w, poff = A["f_first"], A["f_w"], A["f_panel_off"]

        def f(a: int, b: int):
            pa, pb = int(self.pos[a]), int(self.pos[b])
            s = int(self.front_of_pos[pb])
            lcol = int(self.pstart[pb] - first[s])
            ga = int(self.pstart[pa])
            if ga < first[s] + w[s]:
                lrow = ga - int(first[s])
            else:
                k = int(np.searchsorted(self.border_rows[s], ga))
                assert self.border_rows[s][k] == ga
                lrow = int(w[s]) + k
            return int(poff[s]) + lrow * int(w[s]) + lcol, int(w[s]), -1
        return f


def build_front_plan(param_size, ptrs, inds, ordering: str = "auto", tau: float = 0.12, merge_flops: float = 4e4,
                     merge_max_r: int = SMALL_CLASSES[1], small_limit: Optional[int] = None, split_wide: bool = True) -> FrontPlan:
    param_size = np.asarray(param_size, dtype=np.int64)
    ptrs = np.asarray(ptrs, dtype=np.int64)
    inds = np.asarray(inds, dtype=np.int64)
    N = int(param_size.shape[0])
    from .sparse import minimum_degree_order
    cands = {}
    if ordering in ("auto", "nd"):
        cands["nd"] = nested_dissection_order(N, ptrs, inds)
    if ordering in ("auto", "mindeg"):
        cands["mindeg"] = minimum_degree_order(N, ptrs, inds, param_size)
    if ordering == "natural":
        cands["natural"] = np.arange(N, dtype=np.int64)
    if not cands:
        raise ValueError(ordering)
    best = None
    for name, o in cands.items():
        pos0, parent, struct = _column_structures(N, ptrs, inds, o)
        fl = _flops_of(struct, param_size[o])
        if best is None or fl < best[0]:
            best = (fl, name, o, pos0, parent, struct)
    col_flops, oname, order0, pos0, parent, struct = best
    dims0 = param_size[order0]

    # ---- fundamental supernodes ----
    nchild = np.zeros(N, dtype=np.int64)
    for j in range(N):
        if parent[j] >= 0:
            nchild[parent[j]] += 1
    firsts = [0]
    for j in range(1, N):
        if not (parent[j - 1] == j and len(struct[j - 1]) == len(struct[j]) + 1 and nchild[j] == 1):
            firsts.append(j)
    firsts.append(N)
    sn = [(firsts[i], firsts[i + 1]) for i in range(len(firsts) - 1)]
    S0 = len(sn)
    sn_of = np.empty(N, dtype=np.int64)
    for s, (a, b) in enumerate(sn):
        sn_of[a:b] = s
    cols = [list(range(a, b)) for a, b in sn]
    below = [struct[b - 1] for a, b in sn]
    wsz = [int(dims0[a:b].sum()) for a, b in sn]
    bsz = [int(dims0[struct[b - 1]].sum()) if len(struct[b - 1]) else 0 for a, b in sn]
    par = np.array([sn_of[parent[b - 1]] if parent[b - 1] >= 0 else -1 for a, b in sn], dtype=np.int64)
    kids: List[List[int]] = [[] for _ in range(S0)]
    for s in range(S0):
        if par[s] >= 0:
            kids[par[s]].append(s)
    alive = np.ones(S0, dtype=bool)
    base = [_front_cost(wsz[s], bsz[s]) for s in range(S0)]   # flops of the fundamental supernodes merged into s so far
    # ---- relaxed amalgamation, bottom-up (supernode indices are a topological order): a child joins its parent while the dense
    # flops of the merged front stay within (1 + tau) of what its fundamental supernodes cost, or -- small fronts only -- within a
    # fixed allowance (a front costs a CTA, a trip through the update-matrix arena and a slot in every level's launch) ----
    for p in range(S0):
        changed = True
        while changed and kids[p]:
            changed = False
            for c in sorted(kids[p], key=lambda c: (wsz[c] + bsz[c], c)):
                sep = base[c] + base[p]
                mer = _front_cost(wsz[c] + wsz[p], bsz[p])
                r_m = wsz[c] + wsz[p] + bsz[p]
                if mer <= (1.0 + tau) * sep or (mer <= sep + merge_flops and r_m <= merge_max_r):
                    cols[p] = cols[c] + cols[p]
                    wsz[p] += wsz[c]
                    base[p] = sep
                    alive[c] = False
                    kids[p].remove(c)
                    for g in kids[c]:
                        par[g] = p
                        kids[p].append(g)
                    kids[c] = []
                    changed = True
                    break
    # ---- wide supernodes are SPLIT into a chain of fronts that fit the shared-memory kernel (a supernode's columns form a dense
    # trapezoid: piece k = its columns as pivots, the later pieces' columns + the supernode's border as border rows; same flops, one
    # more trip of an update matrix through the arena).  Left whole -- for the dense path -- : borderless fronts wider than
    # SPLIT_ROOT_W (a plain dense Cholesky: the tiled DMMA kernel is the better engine) and fronts whose single columns do not fit. ----
    fr_cols: List[List[int]] = []
    fr_below: List[np.ndarray] = []
    for sidx in (q for q in range(S0) if alive[q]):
        cs, bl = cols[sidx], np.asarray(below[sidx], dtype=np.int64)
        wtot, btot = int(dims0[cs].sum()), (int(dims0[bl].sum()) if len(bl) else 0)
        fits = wtot <= SMALL_MAX_W and small_smem_bytes(wtot, btot, 4) <= SMALL_SMEM_LIMIT
        if fits or not split_wide or (btot == 0 and wtot > SPLIT_ROOT_W):
            fr_cols.append(list(cs)); fr_below.append(bl)
            continue
        pieces, i0 = [], 0
        while i0 < len(cs):
            i1, wacc = i0, 0
            rest = int(dims0[cs[i0:]].sum())
            while i1 < len(cs):
                wn = wacc + int(dims0[cs[i1]])
                if i1 > i0 and (wn > SPLIT_MAX_W or small_smem_bytes(wn, rest - wn + btot, 4) > SMALL_SMEM_LIMIT):
                    break
                wacc, i1 = wn, i1 + 1
            pieces.append((i0, i1))
            i0 = i1
        if any(small_smem_bytes(int(dims0[cs[a:b_]].sum()), int(dims0[cs[b_:]].sum()) + btot, 4) > SMALL_SMEM_LIMIT for a, b_ in pieces):
            fr_cols.append(list(cs)); fr_below.append(bl)       # even single columns do not fit: dense path
            continue
        for a, b_ in pieces:
            fr_cols.append(list(cs[a:b_]))
            fr_below.append(np.concatenate([np.asarray(cs[b_:], dtype=np.int64), bl]))
    S = len(fr_cols)
    keep = list(range(S))
    cols, below = fr_cols, fr_below
    # ---- final elimination order: fronts in topological order, each front's columns contiguous ----
    old_positions = np.array([j for s in keep for j in cols[s]], dtype=np.int64)
    assert sorted(old_positions.tolist()) == list(range(N))
    newpos_of_old = np.empty(N, dtype=np.int64)
    newpos_of_old[old_positions] = np.arange(N)
    order = order0[old_positions]
    pos = np.empty(N, dtype=np.int64)
    pos[order] = np.arange(N)
    dims = param_size[order]
    pstart = np.concatenate([[0], np.cumsum(dims)[:-1]]).astype(np.int64)
    n = int(dims.sum())
    orig_start = np.concatenate([[0], np.cumsum(param_size)[:-1]]).astype(np.int64)
    col_start = orig_start[order]
    perm = np.concatenate([np.arange(col_start[k], col_start[k] + dims[k]) for k in range(N)]).astype(np.int64) if N else np.zeros(0, np.int64)

    f_w = np.zeros(S, dtype=np.int32)
    f_b = np.zeros(S, dtype=np.int32)
    f_first = np.zeros(S, dtype=np.int32)
    f_parent = -np.ones(S, dtype=np.int32)
    front_of_pos = np.empty(N, dtype=np.int64)
    border_rows: List[np.ndarray] = []
    k0 = 0
    for t, s in enumerate(keep):
        nc = len(cols[s])
        front_of_pos[k0:k0 + nc] = t
        f_first[t] = pstart[k0]
        f_w[t] = int(dims[k0:k0 + nc].sum())
        bp = np.sort(newpos_of_old[below[s]]) if len(below[s]) else np.zeros(0, dtype=np.int64)
        assert len(bp) == 0 or bp[0] >= k0 + nc
        rows = np.concatenate([np.arange(pstart[q], pstart[q] + dims[q]) for q in bp]).astype(np.int64) if len(bp) else np.zeros(0, np.int64)
        border_rows.append(rows)
        f_b[t] = rows.shape[0]
        k0 += nc
    for t in range(S):
        if f_b[t] > 0:
            first_row_pos = int(np.searchsorted(pstart, border_rows[t][0], side="right") - 1)
            f_parent[t] = int(front_of_pos[first_row_pos])
            assert f_parent[t] > t
    f_r = f_w + f_b
    # ---- depth (root = 0); children sit exactly one depth below their parent ----
    f_depth = np.zeros(S, dtype=np.int32)
    for t in range(S - 1, -1, -1):
        if f_parent[t] >= 0:
            f_depth[t] = f_depth[f_parent[t]] + 1
    max_depth = int(f_depth.max()) if S else 0
    children: List[List[int]] = [[] for _ in range(S)]
    for t in range(S):
        if f_parent[t] >= 0:
            children[f_parent[t]].append(t)
    # ---- size class: 0..2 shared-memory kernel, 3 dense DMMA kernel ----
    f_class = np.zeros(S, dtype=np.int32)
    for t in range(S):
        r = int(f_r[t])
        too_big = (small_limit is not None and r > small_limit) or int(f_w[t]) > SMALL_MAX_W or \
            small_smem_bytes(int(f_w[t]), int(f_b[t]), len(children[t])) > SMALL_SMEM_LIMIT or len(children[t]) > SMALL_MAX_CHILDREN
        f_class[t] = 3 if too_big else min(int(np.searchsorted(np.array(SMALL_CLASSES), r)), 2)
    # ---- storage: panels, update-matrix arena (by depth parity), border-vector arena ----
    f_panel_off = np.zeros(S, dtype=np.int64)
    off = 0
    for t in range(S):
        f_panel_off[t] = off
        off += int(f_r[t]) * int(f_w[t])
        off += off & 1                      # keep every panel 16-byte aligned
    data_size = off
    f_wpad = np.zeros(S, dtype=np.int32)    # big fronts: pivot columns padded to a multiple of the block-column width
    f_np = np.zeros(S, dtype=np.int32)      # big fronts: padded order of the dense front matrix
    f_cb_off = np.zeros(S, dtype=np.int64)  # offset of CB[0][0] (first border row / column) in the arena of parity depth % 2
    f_cb_ld = np.zeros(S, dtype=np.int32)
    f_fr_off = np.zeros(S, dtype=np.int64)  # big fronts: offset of the dense front matrix F (np x np)
    f_u_off = np.zeros(S, dtype=np.int64)
    arena = [0, 0]
    varena = [0, 0]
    for d in range(max_depth, -1, -1):
        a = 0
        v = 0
        for t in np.nonzero(f_depth == d)[0]:
            if f_class[t] == 3:
                wp = -(-int(f_w[t]) // BIG_TW) * BIG_TW
                npad = -(-(wp + int(f_b[t])) // BIG_TM) * BIG_TM
                f_wpad[t], f_np[t] = wp, npad
                f_fr_off[t] = a
                f_cb_off[t] = a + wp * npad + wp
                f_cb_ld[t] = npad
                a += npad * npad
            else:
                ld = int(f_b[t]) + (int(f_b[t]) & 1)
                f_cb_off[t] = a
                f_cb_ld[t] = ld
                a += int(f_b[t]) * ld
            f_u_off[t] = v
            v += int(f_b[t]) + (int(f_b[t]) & 1)
        arena[d & 1] = max(arena[d & 1], a)
        varena[d & 1] = max(varena[d & 1], v)
    arena_size = max(arena) + 2
    varena_size = max(varena) + 2
    # ---- children lists + relative row maps ----
    child_ptr = np.zeros(S + 1, dtype=np.int32)
    child_list: List[int] = []
    rel_ptr = np.zeros(S + 1, dtype=np.int64)      # per CHILD front: its border rows' local indices in the parent front
    rel_list: List[np.ndarray] = []
    for t in range(S):
        child_list.extend(children[t])
        child_ptr[t + 1] = len(child_list)
    racc = 0
    for t in range(S):
        p = int(f_parent[t])
        if p >= 0:
            g = border_rows[t]
            inp = g < f_first[p] + f_w[p]
            loc = np.where(inp, g - f_first[p], f_w[p] + np.searchsorted(border_rows[p], g))
            if len(border_rows[p]):
                kk = np.clip(loc - f_w[p], 0, len(border_rows[p]) - 1)
                assert bool(np.all(inp | (border_rows[p][kk] == g)))
            else:
                assert bool(np.all(inp))
            assert bool(np.all(g >= f_first[p]))
            rel_list.append(loc.astype(np.int32))
            racc += len(loc)
        rel_ptr[t + 1] = racc
    f_rel = np.concatenate(rel_list).astype(np.int32) if rel_list else np.zeros(0, dtype=np.int32)
    # what the kernels would otherwise binary-search in global memory, per child front t of parent p:
    #   c_jw[t]  = number of t's border rows that are PIVOTS of p (rel < w_p);
    #   c_sp[c_sp_ptr[t] + s] = first border row of t whose image lies at or after border row 32 s of p (stripe pointers)
    c_jw = np.zeros(S, dtype=np.int32)
    c_sp_ptr = np.zeros(S + 1, dtype=np.int64)
    c_sp_list: List[np.ndarray] = []
    acc = 0
    for t in range(S):
        p = int(f_parent[t])
        if p >= 0:
            rel = f_rel[rel_ptr[t]:rel_ptr[t + 1]]
            c_jw[t] = int(np.searchsorted(rel, int(f_w[p])))
            ns = -(-int(f_b[p]) // STRIPE) + 1
            sp = np.searchsorted(rel, int(f_w[p]) + STRIPE * np.arange(ns)).astype(np.int32)
            c_sp_list.append(sp)
            acc += ns
        c_sp_ptr[t + 1] = acc
    c_sp = np.concatenate(c_sp_list).astype(np.int32) if c_sp_list else np.zeros(1, dtype=np.int32)
    #   c_inv[c_inv_ptr[t] + l] = border row of t whose image is row l of p's front, or -1 (the gather kernel's map)
    c_inv_ptr = np.zeros(S + 1, dtype=np.int64)
    c_inv_list: List[np.ndarray] = []
    acc = 0
    for t in range(S):
        p = int(f_parent[t])
        if p >= 0:
            inv = -np.ones(int(f_r[p]), dtype=np.int32)
            rel = f_rel[rel_ptr[t]:rel_ptr[t + 1]]
            inv[rel] = np.arange(rel.shape[0], dtype=np.int32)
            c_inv_list.append(inv)
            acc += inv.shape[0]
        c_inv_ptr[t + 1] = acc
    c_inv = np.concatenate(c_inv_list).astype(np.int32) if c_inv_list else np.zeros(1, dtype=np.int32)
    rows_ptr = np.zeros(S + 1, dtype=np.int64)
    rows_ptr[1:] = np.cumsum(f_b)
    f_rows = np.concatenate(border_rows).astype(np.int32) if S else np.zeros(0, dtype=np.int32)
    # ---- schedule ----
    # a launch's dynamic shared memory is its largest front's: fronts of one (depth, class) are split by occupancy bucket
    # (CTAs per SM that fit), so that a few large fronts do not cap the residency of the many small ones
    def bucket(t):
        if f_class[t] == 3:
            return 0
        sm = small_smem_bytes(int(f_w[t]), int(f_b[t]), len(children[t]))
        return int(np.searchsorted(np.array(SMEM_BUCKETS), sm))
    f_bucket = np.array([bucket(t) for t in range(S)], dtype=np.int32)
    sched = np.array(sorted(range(S), key=lambda t: (-int(f_depth[t]), int(f_class[t]), int(f_bucket[t]), -int(f_r[t]), t)), dtype=np.int32)
    launches = []
    i = 0
    while i < S:
        t = int(sched[i])
        d, c, bk = int(f_depth[t]), int(f_class[t]), int(f_bucket[t])
        j = i + 1
        if c != 3:
            while j < S and int(f_depth[sched[j]]) == d and int(f_class[sched[j]]) == c and int(f_bucket[sched[j]]) == bk:
                j += 1
            smem = max(small_smem_bytes(int(f_w[q]), int(f_b[q]), len(children[q])) for q in sched[i:j])
            launches.append((d, c, i, j - i, smem, max(int(f_r[q]) for q in sched[i:j]), max(int(f_r[q]) * int(f_w[q]) for q in sched[i:j]),
                             0, 0, 0, 0, 0))
        else:   # one front: (.., np, pivot block columns, offset of F in the arena, first pivot [info base], front index)
            launches.append((d, c, i, 1, 0, int(f_np[t]), int(f_wpad[t]) // BIG_TW, int(f_fr_off[t]), int(f_first[t]), t, int(f_w[t]), int(f_b[t])))
        i = j
    launches = np.array(launches, dtype=np.int64).reshape(-1, 12)
    # flat descriptors for the shared-memory kernel (see thb200.h): per front in launch order, per (parent, child) pair
    fd = np.zeros((max(S, 1), 8), dtype=np.int64)
    pc = np.zeros((max(int(child_ptr[S]), 1), 6), dtype=np.int64)
    for q in range(S):
        t = int(sched[q])
        cb, ce = int(child_ptr[t]), int(child_ptr[t + 1])
        fd[q] = (t, int(f_w[t]), int(f_b[t]), int(f_first[t]), int(f_panel_off[t]), int(f_cb_off[t]), int(f_cb_ld[t]), cb | ((ce - cb) << 32))
    for t in range(S):
        for k, c in enumerate(children[t]):
            rel = f_rel[rel_ptr[c]:rel_ptr[c + 1]]
            pc[int(child_ptr[t]) + k] = (int(f_cb_off[c]), int(f_cb_ld[c]) | (int(f_b[c]) << 32), int(rel[0]), int(rel[-1]), int(c_inv_ptr[c]),
                                         int(f_u_off[c]))
    flops = float(sum(_front_cost(float(f_w[t]), float(f_b[t])) for t in range(S)))
    stats = dict(ordering=oname, column_flops=col_flops, flops=flops, nnz_L=float(sum(int(f_r[t]) * int(f_w[t]) for t in range(S))),
                 fronts=float(S), max_front=float(f_r.max()) if S else 0.0, depth=float(max_depth + 1), big_fronts=float((f_class == 3).sum()),
                 launches=float(len(launches)), cb_doubles=float(sum(int(b) * int(b) for b in f_b)), levels=float(max_depth + 1),
                 num_updates=0.0, num_chains=float(S), chain_levels=float(max_depth + 1))
    arrays = dict(f_w=f_w, f_b=f_b, f_first=f_first, f_parent=f_parent, f_depth=f_depth, f_class=f_class, f_panel_off=f_panel_off,
                  f_wpad=f_wpad, f_np=f_np, f_cb_off=f_cb_off, f_cb_ld=f_cb_ld, f_fr_off=f_fr_off, f_u_off=f_u_off,
                  child_ptr=child_ptr, child_list=np.array(child_list, dtype=np.int32), rel_ptr=rel_ptr, f_rel=f_rel,
                  rows_ptr=rows_ptr, f_rows=f_rows, sched=sched, perm=perm.astype(np.int32), c_jw=c_jw, c_sp_ptr=c_sp_ptr, c_sp=c_sp,
                  c_inv_ptr=c_inv_ptr, c_inv=c_inv, fd=fd.reshape(-1), pc=pc.reshape(-1),
                  pmap=-np.ones(max(int(data_size), 1), dtype=np.int32))
    return FrontPlan(N=N, n=n, param_size=param_size, order=order, pos=pos, dims=dims, pstart=pstart, col_start=col_start, perm=perm, S=S,
                     arrays=arrays, launches=launches, data_size=int(data_size), arena_size=int(arena_size), varena_size=int(varena_size),
                     stats=stats, front_of_pos=front_of_pos, border_rows=border_rows)


# ------------------------------------------------------------------------------------------------ numpy interpreter (tests)
def execute_numpy(plan: FrontPlan, panels: np.ndarray, rhs: np.ndarray):
    """Runs the plan's arrays like the kernels do, for ONE item: `panels` [data_size] holds AtA (+ damping) scattered by
    gram_out_offsets; rhs [n] in ORIGINAL column order.  Returns (x [n] original order, factored panels)."""
    A = plan.arrays
    P = panels.astype(np.float64).copy()
    S = plan.S
    cb: Dict[int, np.ndarray] = {}
    w, b, first, poff = A["f_w"], A["f_b"], A["f_first"], A["f_panel_off"]

    def rel_of(c):
        return A["f_rel"][A["rel_ptr"][c]:A["rel_ptr"][c + 1]]

    def kids(t):
        return A["child_list"][A["child_ptr"][t]:A["child_ptr"][t + 1]]
    for (d, cls, s0, cnt) in plan.launches[:, :4]:
        for t in A["sched"][s0:s0 + cnt]:
            wt, bt = int(w[t]), int(b[t])
            r = wt + bt
            F = np.zeros((r, r))
            F[:, :wt] = P[poff[t]:poff[t] + r * wt].reshape(r, wt)
            for c in kids(t):
                rel = rel_of(c)
                C = np.tril(cb.pop(int(c)))
                F[np.ix_(rel, rel)] += C
            F = np.tril(F)
            L11 = np.linalg.cholesky(F[:wt, :wt] + np.tril(F[:wt, :wt], -1).T)
            F[:wt, :wt] = L11
            if bt:
                L21 = np.linalg.solve(L11, F[wt:, :wt].T).T
                F[wt:, :wt] = L21
                cb[int(t)] = F[wt:, wt:] + np.tril(F[wt:, wt:], -1).T - L21 @ L21.T
            P[poff[t]:poff[t] + r * wt] = F[:, :wt].reshape(-1)
    y = rhs[plan.perm].astype(np.float64).copy()
    ub: Dict[int, np.ndarray] = {}
    for (d, cls, s0, cnt) in plan.launches[:, :4]:   # forward, deepest first
        for t in A["sched"][s0:s0 + cnt]:
            wt, bt = int(w[t]), int(b[t])
            r = wt + bt
            Lp = P[poff[t]:poff[t] + r * wt].reshape(r, wt)
            u = np.zeros(r)
            u[:wt] = y[first[t]:first[t] + wt]
            for c in kids(t):
                u[rel_of(c)] += ub.pop(int(c))
            yt = np.linalg.solve(np.tril(Lp[:wt]), u[:wt])
            y[first[t]:first[t] + wt] = yt
            if bt:
                ub[int(t)] = u[wt:] - Lp[wt:] @ yt
    x = y
    for (d, cls, s0, cnt) in plan.launches[::-1, :4]:    # backward, root first
        for t in A["sched"][s0:s0 + cnt]:
            wt, bt = int(w[t]), int(b[t])
            r = wt + bt
            Lp = P[poff[t]:poff[t] + r * wt].reshape(r, wt)
            rows = A["f_rows"][A["rows_ptr"][t]:A["rows_ptr"][t + 1]]
            tt = x[first[t]:first[t] + wt] - (Lp[wt:].T @ x[rows] if bt else 0.0)
            x[first[t]:first[t] + wt] = np.linalg.solve(np.tril(Lp[:wt]).T, tt)
    out = np.empty(plan.n)
    out[plan.perm] = x
    return out, P
