# Times the LM loop with the block-sparse solver on C2-size and C5-size (sphere2500-like) pose graphs.
import sys, os, time, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theseus_b200 as th
from theseus_b200.datasets import pose_graph_synthetic_3d, pose_graph_sphere, build_pose_graph_objective
which = sys.argv[1] if len(sys.argv) > 1 else "c5"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
iters = 10
layout = sys.argv[3] if len(sys.argv) > 3 else None          # lane | item | lane_root | lane_tiled | lane_tiled_root
supernodal = "supernodal" in sys.argv[4:]                      # chain-piece substitution kernels
t0 = time.time()
data = pose_graph_sphere(50, 50, B) if which == "c5" else pose_graph_synthetic_3d(256, B)
print("data", round(time.time() - t0, 1), "s", "edges", len(data["edges"]), flush=True)
t0 = time.time()
objective, poses = build_pose_graph_objective(th, data, torch.device("cuda", 0))
opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                            max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0, linear_solver_kwargs=dict(layout=layout, supernodal_solve=supernodal),
                            cuda_graph=("graph" in sys.argv[4:] or os.environ.get("BENCH_GRAPH") == "1"))
print("objective+symbolic", round(time.time() - t0, 1), "s", opt.linear_solver.symbolic_stats, flush=True)
kw = dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
inputs = {p.name: data["poses"][i].cuda() for i, p in enumerate(poses)}
layer = th.TheseusLayer(opt)
def step():
    with torch.no_grad():
        return layer.forward(inputs, optimizer_kwargs=kw)
t0 = time.time(); values, info = step(); torch.cuda.synchronize(); print("first step (incl. engine build)", round(time.time() - t0, 2), "s")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps): values, info = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps(dict(config=which, B=B, ms_per_step=ms, lm_it_per_s=iters * 1e3 / ms, err0=float(info.err_history[:, 0].mean()) if info.err_history is not None else None,
                      final_err=float(info.last_err.mean()))))
# phase breakdown of one iteration
lin = opt.linear_solver.linearization
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
print("linearize ms", t(lin.linearize))
lam = torch.full((B,), 1e-3, dtype=torch.float64, device="cuda")
print("solve (gram+damp+factor+solve) ms", t(lambda: opt.linear_solver.solve(damping=lam, ellipsoidal_damping=True, damping_eps=1e-8)))
eng = objective.engine()
print("error_metric ms", t(lambda: eng.error_metric("cur")))
# numeric phase / substitutions separately (the solver's backend protocol)
ls = opt.linear_solver
A64, b64 = lin.A_val.double().contiguous(), lin.b.double().contiguous()
from theseus_b200.optimizer import convert_to_alpha_beta_damping_tensors
alpha, beta = convert_to_alpha_beta_damping_tensors(lam, 1e-8, True, B, A64.device, torch.float64)
Atb = ls._numeric(A64, b64, alpha, beta)
print("numeric (zero-fill + gram + damp + factor + Atb) ms", t(lambda: ls._numeric(A64, b64, alpha, beta)))
print("substitutions ms", t(lambda: ls._substitute(Atb)))
