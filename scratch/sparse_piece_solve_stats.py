# Launch counts of the supernodal (chain-piece) substitution schedule (sparse.piece_solve_lists) on the C5 plan.
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theseus_b200.datasets import pose_graph_sphere
from theseus_b200.sparse import analyze, piece_solve_lists, root_split
data = pose_graph_sphere(50, 50, 1)
N = len(data["poses"]); adj = [set([i]) for i in range(N)]
for e in data["edges"]:
    i, j = int(e[0]), int(e[1]); adj[i].add(j); adj[j].add(i)
ptrs, inds = [0], []
for i in range(N):
    inds += sorted(adj[i]); ptrs.append(len(inds))
P = analyze(np.full(N, 6), np.array(ptrs), np.array(inds))
sp = root_split(P)
print(f"columns {P.N}, column levels {int(P.level.max()) + 1}; root cut {sp['cut']} (column levels below the cut: {int(P.level[sp['cut']])})")
for W in (1, 2, 4, 8):
    for cut in (None, sp["cut"]):
        ps = piece_solve_lists(P, max_width=W, cut=cut)
        A = P.arrays
        ext = int(sum(ps["fr_ext_end"][j] - A["fr_ptr"][j] for j in range(ps["cut"])))
        tot = int(A["fr_ptr"][ps["cut"]])
        # backward: external block loads are per column, the x_i loads per piece
        xi_loads_piece = int(sum(A["bc_ptr"][f + w] - ps["bc_int_end"][f + w - 1] for f, w in zip(ps["first"], ps["width"])))
        xi_loads_col = int(sum(A["bc_ptr"][j + 1] - ps["bc_int_end"][j] for j in range(ps["cut"])))
        print(f"width {W}, cut {cut}: pieces {len(ps['first'])}, piece levels {int(ps['level'].max()) + 1}, launches {len(ps['launches'])}; "
              f"forward blocks external {ext} of {tot}; backward x_i loads {xi_loads_piece} (per column today: {xi_loads_col})")
