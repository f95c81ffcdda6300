# Prototype: nested dissection + supernodes on the C5 block graph; census of fronts (flops / sizes).
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import deque

def c5_graph(rings=50, per=50):
    N = rings * per
    edges = [(i, i + 1) for i in range(N - 1)] + [(i, i + per) for i in range(N - per)]
    adj = [set() for _ in range(N)]
    for a, b in edges:
        adj[a].add(b); adj[b].add(a)
    return N, [sorted(a) for a in adj]

def bfs_levels(adj, start, mask, stamp, mark):
    # returns list of levels (lists) within nodes where mask[node]==stamp
    mark[start] = True
    levels = [[start]]
    seen = [start]
    while True:
        nxt = []
        for u in levels[-1]:
            for v in adj[u]:
                if mask[v] == stamp and not mark[v]:
                    mark[v] = True; nxt.append(v); seen.append(v)
        if not nxt: break
        levels.append(nxt)
    for v in seen: mark[v] = False
    return levels, seen

def nd_order(N, adj, leaf=8):
    mask = np.zeros(N, dtype=np.int64)   # component stamp
    mark = np.zeros(N, dtype=bool)
    order = []
    next_stamp = [1]
    def rec(nodes):
        # nodes: list; all have mask == s
        s = next_stamp[0]; next_stamp[0] += 1
        for v in nodes: mask[v] = s
        if len(nodes) <= leaf:
            # local min-degree-ish: order by degree within set
            order.extend(sorted(nodes, key=lambda v: (sum(1 for x in adj[v] if mask[x] == s), v)))
            for v in nodes: mask[v] = 0
            return
        # components
        remaining = set(nodes)
        comps = []
        for v in nodes:
            if v in remaining:
                lv, seen = bfs_levels(adj, v, mask, s, mark)
                comps.append(seen)
                remaining.difference_update(seen)
        if len(comps) > 1:
            for c in comps: rec(c)
            return
        # pseudo-peripheral node
        start = nodes[0]
        levels, _ = bfs_levels(adj, start, mask, s, mark)
        for _ in range(4):
            last = levels[-1]
            cand = min(last, key=lambda v: (sum(1 for x in adj[v] if mask[x] == s), v))
            l2, _ = bfs_levels(adj, cand, mask, s, mark)
            if len(l2) > len(levels):
                levels = l2; start = cand
            else:
                break
        if len(levels) < 3:
            order.extend(sorted(nodes)); 
            for v in nodes: mask[v] = 0
            return
        tot = len(nodes)
        cum = np.cumsum([len(l) for l in levels])
        best, bi = None, None
        for i in range(1, len(levels) - 1):
            left = cum[i - 1]; right = tot - cum[i]
            bal = min(left, right) / max(left, right, 1)
            if bal < 0.4: continue
            score = (len(levels[i]), -bal)
            if best is None or score < best: best, bi = score, i
        if bi is None:
            bi = int(np.argmin(np.abs(cum - tot / 2)));  bi = min(max(bi, 1), len(levels) - 2)
        sep = levels[bi]
        left = [v for l in levels[:bi] for v in l]
        right = [v for l in levels[bi + 1:] for v in l]
        # refine: separator node not adjacent to right side can move to left
        rs = set(right)
        sep2, moved = [], []
        for v in sep:
            if any((x in rs) for x in adj[v]): sep2.append(v)
            else: moved.append(v)
        left += moved; sep = sep2
        for v in sep: mask[v] = 0
        rec(left); rec(right)
        order.extend(sorted(sep))
    rec(list(range(N)))
    return np.array(order, dtype=np.int64)

def symbolic(N, adj, order):
    pos = np.empty(N, dtype=np.int64); pos[order] = np.arange(N)
    # column structures by elimination with parent merging (standard): struct[j] = (adj_lower(j) U union children structs) \ {j}
    parent = -np.ones(N, dtype=np.int64)
    struct = [None] * N
    children = [[] for _ in range(N)]
    for j in range(N):
        v = order[j]
        s = set(int(pos[x]) for x in adj[v] if pos[x] > j)
        for c in children[j]:
            s.update(struct[c]); 
        s.discard(j)
        struct[j] = s
        if s:
            p = min(s); parent[j] = p; children[p].append(j)
    struct = [np.array(sorted(s), dtype=np.int64) for s in struct]
    return pos, parent, struct

def supernodes(N, parent, struct, dims, relax_zero_frac=0.0, max_w=10**9):
    # fundamental supernodes: j+1 == parent[j], struct[j] == {j+1} U struct[j+1], j only child of j+1 ... (only child not required for supernode validity)
    nchild = np.zeros(N, dtype=np.int64)
    for j in range(N):
        if parent[j] >= 0: nchild[parent[j]] += 1
    first = [0]
    for j in range(1, N):
        ok = parent[j - 1] == j and len(struct[j - 1]) == len(struct[j]) + 1 and nchild[j] == 1
        if not ok: first.append(j)
    first.append(N)
    sn = [(first[i], first[i + 1]) for i in range(len(first) - 1)]
    return sn

def census(N, adj, order, d=6, label=""):
    t0 = time.time()
    pos, parent, struct = symbolic(N, adj, order)
    nnzL = sum((len(s) + 1) * d * d for s in struct) - N * (d * d - d * (d + 1) // 2)
    fl = 0.0
    for j in range(N):
        r = len(struct[j]) * d
        # scalar columns of block j
        for c in range(d):
            rr = r + (d - 1 - c)
            fl += rr * rr + 2 * rr  # approx flops (mul+add) of right-looking col
    sn = supernodes(N, parent, struct, None)
    print(f"{label}: nnzL={nnzL/1e6:.3f}M flops={fl/1e9:.3f}G supernodes={len(sn)} maxfront={max(len(s) for s in struct)+1} ({time.time()-t0:.1f}s)")
    return pos, parent, struct, sn

if __name__ == "__main__":
    N, adj = c5_graph()
    from theseus_b200.sparse import minimum_degree_order
    ptrs = np.zeros(N + 1, dtype=np.int64); inds = []
    for i in range(N):
        row = sorted(adj[i] + [i]); inds.extend(row); ptrs[i + 1] = len(inds)
    inds = np.array(inds)
    t0 = time.time(); md = minimum_degree_order(N, ptrs, inds, np.full(N, 6)); print("mindeg", time.time() - t0)
    census(N, adj, md, label="mindeg")
    for leaf in (4, 8, 16, 32):
        t0 = time.time(); o = nd_order(N, adj, leaf=leaf); print("nd", leaf, time.time() - t0)
        assert sorted(o.tolist()) == list(range(N))
        census(N, adj, o, label=f"nd leaf={leaf}")

def front_cost(w, b):
    return w ** 3 / 3.0 + w * w * b + w * b * b

def amalgamate(N, parent, struct, d, tau=0.15, abs_allow=0.0, small_r=0):
    """supernode tree -> merged fronts.  Returns list of (cols list (elimination positions), below-rows array)."""
    sn = supernodes(N, parent, struct, None)
    S = len(sn)
    sn_of = np.empty(N, dtype=np.int64)
    for s, (a, b) in enumerate(sn): sn_of[a:b] = s
    cols = [list(range(a, b)) for a, b in sn]
    below = [struct[b - 1] for a, b in sn]          # rows below the supernode (positions)
    par = np.array([sn_of[parent[b - 1]] if parent[b - 1] >= 0 else -1 for a, b in sn])
    kids = [[] for _ in range(S)]
    for s in range(S):
        if par[s] >= 0: kids[par[s]].append(s)
    alive = np.ones(S, dtype=bool)
    # bottom-up (supernodes are in postorder-ish increasing column order => children before parents)
    for p in range(S):
        if not kids[p]: continue
        changed = True
        while changed:
            changed = False
            ks = sorted([c for c in kids[p] if alive[c]], key=lambda c: len(cols[c]) + len(below[c]))
            for c in ks:
                wc, bc = len(cols[c]) * d, len(below[c]) * d
                wp, bp = len(cols[p]) * d, len(below[p]) * d
                sep = front_cost(wc, bc) + front_cost(wp, bp)
                mer = front_cost(wc + wp, bp)
                if mer <= (1 + tau) * sep + abs_allow or (wc + wp + bp) <= small_r:
                    cols[p] = cols[c] + cols[p]
                    alive[c] = False
                    kids[p].remove(c)
                    for g in kids[c]:
                        par[g] = p; kids[p].append(g)
                    kids[c] = []
                    changed = True
                    break
    out = [(cols[s], below[s], int(par[s])) for s in range(S) if alive[s]]
    return out

def front_census(fronts, d=6, label=""):
    tot = 0.0; nnz = 0
    buckets = {}
    for cols, below, _ in fronts:
        w, b = len(cols) * d, len(below) * d
        f = front_cost(w, b) ; tot += f
        nnz += w * (w + b)
        key = min((w + b) // 48 * 48, 960)
        e = buckets.setdefault(key, [0, 0.0, 0])
        e[0] += 1; e[1] += f; e[2] = max(e[2], w)
    print(f"{label}: fronts={len(fronts)} dense-flops={2*tot/1e9:.3f}G (x2 for mul+add... cost in FMA {tot/1e9:.3f}G) panel-nnz={nnz/1e6:.3f}M")
    for k in sorted(buckets):
        e = buckets[k]
        print(f"   r in [{k},{k+48}): n={e[0]:5d} flops_share={e[1]/tot:.3f} max_w={e[2]}")

if __name__ == "__main__":
    o = nd_order(N, adj, leaf=8)
    pos, parent, struct, sn = census(N, adj, o, label="nd8")
    for tau, aa, sr in ((0.0, 0, 0), (0.1, 0, 0), (0.2, 2e4, 0), (0.3, 5e4, 48), (0.5, 1e5, 64)):
        fr = amalgamate(N, parent, struct, 6, tau, aa, sr)
        front_census(fr, label=f"tau={tau} abs={aa} small_r={sr}")

def cb_census(fronts, d=6, label=""):
    sb = sum((len(b) * d) ** 2 for c, b, p in fronts)
    sp = sum((len(b) * d + len(c) * d) * len(c) * d for c, b, p in fronts)
    # depth
    S = len(fronts)
    print(f"{label}: sum b^2 = {sb/1e6:.2f}M doubles (lower half {sb/2e6:.2f}M), panels {sp/1e6:.2f}M")

if __name__ == "__main__":
    for tau, aa, sr in ((0.0, 0, 0), (0.1, 0, 0), (0.15, 0, 60), (0.2, 0, 96), (0.25, 0, 128)):
        fr = amalgamate(N, parent, struct, 6, tau, aa, sr)
        front_census(fr, label=f"tau={tau} abs={aa} small_r={sr}")
        cb_census(fr, label="   CB")
