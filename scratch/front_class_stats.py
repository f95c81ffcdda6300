# Per-launch-class arithmetic of the C5 front plan: fronts, flops, panel bytes, shared memory; tells where a class is far from its bound.
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theseus_b200.frontal import build_front_plan, small_smem_bytes
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_frontal import _graph_csr
N = 2500
e = [(i, i + 1) for i in range(N - 1)] + [(i, i + 50) for i in range(N - 50)]
ptrs, inds = _graph_csr(N, e)
plan = build_front_plan(np.full(N, 6), ptrs, inds)
A = plan.arrays
w, b, cls, dep = A["f_w"], A["f_b"], A["f_class"], A["f_depth"]
nch = np.diff(A["child_ptr"])
def fl(w, b):
    return w**3 / 3 + w * w * b + w * b * b
def thr(smem, c):
    if c == 0: return 64
    if c == 1: return 128
    if c == 3: return 0
    return 256 if smem <= 56 * 1024 else (512 if smem <= 113 * 1024 else 1024)
rows = {}
for t in range(len(w)):
    sm = small_smem_bytes(int(w[t]), int(b[t]), int(nch[t])) if cls[t] < 3 else 0
    k = thr(sm, int(cls[t]))
    r = rows.setdefault(k, dict(n=0, flops=0.0, panel=0, cb=0, wmax=0, rmax=0, depths=set()))
    r["n"] += 1; r["flops"] += fl(float(w[t]), float(b[t])); r["panel"] += int(w[t] + b[t]) * int(w[t]); r["cb"] += int(b[t]) ** 2 // 2
    r["wmax"] = max(r["wmax"], int(w[t])); r["rmax"] = max(r["rmax"], int(w[t] + b[t])); r["depths"].add(int(dep[t]))
for k in sorted(rows):
    r = rows[k]
    print(f"threads {k:5d}: fronts {r['n']:5d} MFLOP/item {r['flops']/1e6:8.2f} panel KB/item {r['panel']*8/1e3:8.1f} cb KB/item {r['cb']*8/1e3:8.1f} wmax {r['wmax']} rmax {r['rmax']} depths {sorted(r['depths'])}")
L = plan.launches
for row in L:
    d, c, bgn, cnt, smem = row[:5]
    ts = A["sched"][bgn:bgn + cnt]
    print(f"depth {d:3d} cls {c} count {cnt:4d} smem {smem:7d} thr {thr(smem, c):5d} flops/item {sum(fl(float(w[t]), float(b[t])) for t in ts)/1e6:8.3f}M  w [{w[ts].min()},{w[ts].max()}] b [{b[ts].min()},{b[ts].max()}] nch max {nch[ts].max()}")
