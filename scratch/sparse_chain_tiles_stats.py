# Block loads per update of the chain-tiled schedule (sparse.chain_tiles) on the C5 plan.
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theseus_b200.datasets import pose_graph_sphere
from theseus_b200.sparse import analyze, chain_tiles
data = pose_graph_sphere(50, 50, 1)
N = len(data["poses"]); adj = [set([i]) for i in range(N)]
for e in data["edges"]:
    i, j = int(e[0]), int(e[1]); adj[i].add(j); adj[j].add(i)
ptrs, inds = [0], []
for i in range(N):
    inds += sorted(adj[i]); ptrs.append(len(inds))
P = analyze(np.full(N, 6), np.array(ptrs), np.array(inds))
for W, TR in ((1, 4), (2, 4), (4, 4), (4, 8), (8, 8)):
    t0 = time.time()
    ct = chain_tiles(P, max_width=W, tile_rows=TR)
    n_int = sum(p1 - p0 for L in ct["per_level"] for (_, p0, p1) in L["u_int"])
    ntiles = sum(len(L["tiles"]) for L in ct["per_level"]); nsteps = sum(len(t["steps"]) for L in ct["per_level"] for t in L["tiles"])
    print(f"W={W} TR={TR}: external updates {ct['updates']} ({ct['updates']/P.stats['num_updates']:.2f} of all), internal {n_int}, "
          f"block loads {ct['block_loads']} -> {ct['block_loads']/max(ct['updates'],1):.3f} per external update; tiles {ntiles}, steps {nsteps}  [{time.time()-t0:.1f}s]")
