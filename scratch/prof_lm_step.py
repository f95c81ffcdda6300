# Where do the ~6 ms per LM iteration that are neither linearize nor the linear solve go?  torch profiler kernel table of one C5 step.
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theseus_b200 as th
from theseus_b200.datasets import pose_graph_sphere, build_pose_graph_objective
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
data = pose_graph_sphere(50, 50, B)
objective, poses = build_pose_graph_objective(th, data, torch.device("cuda", 0))
opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                            max_iterations=10, abs_err_tolerance=0, rel_err_tolerance=0, linear_solver_kwargs=dict(layout="front"))
kw = dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
inputs = {p.name: data["poses"][i].cuda() for i, p in enumerate(poses)}
layer = th.TheseusLayer(opt)
def step():
    with torch.no_grad():
        return layer.forward(inputs, optimizer_kwargs=kw)
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
print("one step (10 iterations) ms", e0.elapsed_time(e1))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages()
rows = [(k.key, k.count, getattr(k, "device_time_total", getattr(k, "cuda_time_total", 0.0))) for k in ka if getattr(k, "device_time_total", getattr(k, "cuda_time_total", 0.0)) > 0]
dev = [(n, c, t) for n, c, t in rows if not n.startswith("aten::") and not n.startswith("cuda") and "Memcpy" not in n and "Memset" not in n]
tot = sum(t for _, _, t in dev)
print("GPU kernel time in the step: %.2f ms over %d kernel kinds" % (tot / 1e3, len(dev)))
for n, c, t in sorted(dev, key=lambda r: -r[2])[:40]:
    print("%9.3f ms  n=%5d  %s" % (t / 1e3, c, n[:110]))
mem = [(n, c, t) for n, c, t in rows if "Memcpy" in n or "Memset" in n]
for n, c, t in sorted(mem, key=lambda r: -r[2])[:8]:
    print("%9.3f ms  n=%5d  %s" % (t / 1e3, c, n[:110]))
# GPU idle gaps: sort device events by start and sum the holes
evs = sorted([(e.time_range.start, e.time_range.end) for e in prof.events() if str(e.device_type).endswith("CUDA")], key=lambda r: r[0])
if evs:
    busy, cur_s, cur_e = 0.0, evs[0][0], evs[0][1]
    for s, e in evs[1:]:
        if s > cur_e: busy += cur_e - cur_s; cur_s, cur_e = s, e
        else: cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print("device busy %.2f ms of %.2f ms span" % (busy / 1e3, (evs[-1][1] - evs[0][0]) / 1e3))
