# Config C3 (bundle adjustment, 50 cameras x 1000 points x 8000 reprojections, Huber, batch 32): LM iterations/s with the block-sparse
# solver (lane layout) and with the dense solver.  Scene = the reference-generated fixture tests/golden/ba_c3_huber.npz, batch tiled.
import sys, os, json, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import theseus_b200 as th
from helpers import load, ba_objective
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
solver = sys.argv[2] if len(sys.argv) > 2 else "lane"
g0 = load("ba_c3_huber")
g = {k: g0[k] for k in g0.files}
reps = (B + 1) // 2
for k in ("cam_pose0", "pts0"):
    g[k] = np.tile(g0[k], (1, reps) + (1,) * (g0[k].ndim - 2))[:, :B]
for k in ("known_pose",):
    pass
class G(dict):
    files = list(g.keys())
g = G(g)
t0 = time.time()
objective, cams, pts = ba_objective(th, g)
iters = 10
skw = dict(linear_solver_cls=th.CholeskyDenseSolver) if solver == "dense" else dict(
    linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization, linear_solver_kwargs=dict(layout=solver))
opt = th.LevenbergMarquardt(objective, max_iterations=iters, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0,
                            cuda_graph=len(sys.argv) > 3 and sys.argv[3] == "graph", **skw)
print("objective + symbolic", round(time.time() - t0, 2), "s", getattr(opt.linear_solver, "symbolic_stats", None), flush=True)
kw = dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
inputs = {n: v.tensor.clone() for n, v in objective.optim_vars.items()}
layer = th.TheseusLayer(opt)
def step():
    with torch.no_grad():
        return layer.forward(inputs, optimizer_kwargs=kw)
t0 = time.time(); values, info = step(); torch.cuda.synchronize(); print("first step", round(time.time() - t0, 2), "s")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): values, info = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(json.dumps(dict(config="c3", B=B, solver=solver, ms_per_step=ms, lm_it_per_s=iters * 1e3 / ms, final_err=float(info.last_err.mean()))))
