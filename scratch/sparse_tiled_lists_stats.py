# Launch list / work split of the `lane_tiled` layout (sparse.tile_lane_lists) on the C5 plan, against the plain lane lists.
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theseus_b200.datasets import pose_graph_sphere
from theseus_b200.sparse import analyze, tile_lane_lists
data = pose_graph_sphere(50, 50, 1)
N = len(data["poses"]); adj = [set([i]) for i in range(N)]
for e in data["edges"]:
    i, j = int(e[0]), int(e[1]); adj[i].add(j); adj[j].add(i)
ptrs, inds = [0], []
for i in range(N):
    inds += sorted(adj[i]); ptrs.append(len(inds))
P = analyze(np.full(N, 6), np.array(ptrs), np.array(inds))
t0 = time.time()
lane, tiles = tile_lane_lists(P)
dt = time.time() - t0
T = tiles["tile_tgt"].shape[0]; S = tiles["step_src"].shape[0]
steps = np.diff(tiles["step_ptr"])
n_u = int((lane["u_p1"] - lane["u_p0"]).sum()); total = len(P.arrays["up_a"])
loads = int((tiles["step_src"] >= 0).sum())
kinds = lane["launches"][:, 0]
print(f"tile_lane_lists: {dt:.1f}s; tiles {T}, k steps {S} (per tile: median {int(np.median(steps))}, max {int(steps.max())}); "
      f"staged block loads {loads} ({loads/(total-n_u):.3f} per tiled update)")
print(f"update pairs: {total - n_u} in tiles ({(total-n_u)/total:.3f}), {n_u} per block (U/UH)")
print(f"launches: {len(kinds)} (TU {int((kinds==4).sum())}, U {int((kinds==0).sum())}, UH {int((kinds==3).sum())}, T {int((kinds==1).sum())}, S {int((kinds==2).sum())}) "
      f"vs plain lane {len(P.lane['launches'])}")
heavy = lane["launches"][kinds == 3]
print("largest remaining per-block pair list:", int((lane["u_p1"] - lane["u_p0"]).max()) if len(lane["u_p0"]) else 0)

# the same with the dense root split off (layout lane_tiled_root): the root's assembly is ONE tile launch
from theseus_b200.sparse import root_split
sp = root_split(P)
lane, tiles = tile_lane_lists(P, sp)
kinds = lane["launches"][:, 0]
last_tu = [l for l in lane["launches"] if l[0] == 4][-1]
steps = np.diff(tiles["step_ptr"])
asm = steps[last_tu[3]:last_tu[4]]
print(f"with root split at column {sp['cut']} (root {sp['root_dof']} dof): launches {len(kinds)} (TU {int((kinds==4).sum())}, U {int((kinds==0).sum())}, "
      f"UH {int((kinds==3).sum())}, T {int((kinds==1).sum())}, S {int((kinds==2).sum())}); tiles {tiles['tile_tgt'].shape[0]}, k steps {tiles['step_src'].shape[0]}; "
      f"bottom tiles: max {int(steps[:last_tu[3]].max())} steps; root assembly launch: {len(asm)} tiles, {int(asm.sum())} steps, max {int(asm.max())} per tile")
