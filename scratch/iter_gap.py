# Where does the per-iteration time that is not kernel time go?  Host timeline of one LM step of the bench problem.
import sys, os, time, warnings, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
device = torch.device("cuda", 0)
th, data, objective, poses = bench.build_problem(0, device)
opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=bench.LM_ITERS, step_size=1.0,
                            abs_err_tolerance=0, rel_err_tolerance=0)
layer = th.TheseusLayer(opt)
dev_inputs = {p.name: data["poses"][i].to(device) for i, p in enumerate(poses)}
def step():
    with torch.no_grad():
        return layer.forward(dev_inputs, optimizer_kwargs=bench.LM_KW)
for _ in range(2): step()
torch.cuda.synchronize()
# 1. sync points
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    step()
torch.cuda.set_sync_debug_mode("default")
from collections import Counter
c = Counter()
for x in w:
    c[f"{os.path.basename(x.filename)}:{x.lineno}"] += 1
print("sync points in one step (10 iterations):", dict(c))
# 2. host time of the pieces with the GPU kept out of the way (sync before and after each piece)
lin = opt.linear_solver.linearization
eng = objective.engine()
def host_ms(fn, n=20):
    torch.cuda.synchronize(); t = 0.0
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t += time.perf_counter() - t0; torch.cuda.synchronize()
    return t / n * 1e3
print("host launch time (ms): linearize", host_ms(lin.linearize), " compute_delta", host_ms(lambda: opt.compute_delta(**{k: v for k, v in bench.LM_KW.items()})),
      " error_metric", host_ms(lambda: eng.error_metric("cur")))
t0 = time.perf_counter(); step(); torch.cuda.synchronize(); print("one step wall ms", (time.perf_counter() - t0) * 1e3)
# 3. torch profiler: GPU idle gaps
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"GPU busy {busy/1e3:.2f} ms of span {span/1e3:.2f} ms; gaps > 50 us:")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 50: gaps.append((g, a.name[:40], b.name[:40]))
from collections import defaultdict
agg = defaultdict(lambda: [0, 0.0])
for g, a, b in gaps:
    agg[(a, b)][0] += 1; agg[(a, b)][1] += g
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
    print(f"  {v[0]:3d} x avg {v[1]/v[0]:8.1f} us   after [{k[0]}] before [{k[1]}]")
