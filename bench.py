#!/usr/bin/env python
"""bench.py -- LM iterations/sec on the batched SE3 pose-graph (BASELINE.json config C2) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

A "step" = one pass of the hot path over one batch: one LM solve (max_iterations LM iterations of
linearize -> damped dense Cholesky solve -> retract -> error -> accept/reject) of B=256 pose graphs with 256 SE3
poses each (n=1536 columns), built exactly like examples/pose_graph/pose_graph_cube.py:56-83, fp64, fixed iteration
count (abs/rel tolerance 0) with LM kwargs damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True.
value = LM iterations per second for the whole job (all ranks' batches advance one iteration together).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (dominant kernel = the DMMA Cholesky),
cpu_baseline (oracle port on the host cores, bounded sample), e2e (host buffers -> public API -> host result), and
`sparse_c5` -- a SECOND, separately timed workload reported beside the headline (never mixed into `value`): BASELINE.json's
config C5 (2 500-pose sphere pose graph, batch 512 per GPU = 4096 on 8 GPUs, LM + block-sparse Cholesky), same barrier /
CUDA-event / max-over-ranks timing, 1 warm-up + 2 timed solves; `--no-c5` skips it, a failure inside it is reported as
`sparse_c5.error` and leaves the headline untouched.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NUM_POSES = 256
BATCH = 256
LM_ITERS = 10
LM_KW = dict(damping=1e-3, adaptive_damping=True, ellipsoidal_damping=True)
METRIC = "LM iterations/sec on batched SE3 pose-graph (256 poses, batch 256/GPU, LM + dense Cholesky, fp64)"
UNIT = "LM iterations/s (one iteration = one LM step of a 256-problem batch; aggregate over GPUs)"
WORKLOAD = "C2: synthetic SE3 pose-graph (pose_graph_cube shape: 256 poses, loop_closure_ratio 0.2), batch=256 per GPU, LM(10 it, adaptive+ellipsoidal damping) + CholeskyDenseSolver"


def _measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0), "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    samples=len(sm))


def build_problem(rank, device):
    import theseus_b200 as th
    from theseus_b200.datasets import build_pose_graph_objective, pose_graph_synthetic_3d
    data = pose_graph_synthetic_3d(NUM_POSES, BATCH, translation_noise=0.05, rotation_noise=0.02, loop_closure_ratio=0.2, seed=rank)
    objective, poses = build_pose_graph_objective(th, data, device)
    return th, data, objective, poses


def oracle_spec(data, sl):
    """The same workload as a numpy problem description for the oracle port (CPU baseline only)."""
    P, M = data["poses"][:, sl].numpy(), data["meas"][:, sl].numpy()
    spec = dict(dtype=np.dtype(np.float64), vars=[], costs=[])
    for i in range(P.shape[0]):
        spec["vars"].append(dict(kind="SE3", dof=6, value=P[i]))
    w = data["info"].numpy().reshape(1, 6)
    for e, (i, j) in enumerate(data["edges"]):
        spec["costs"].append(dict(kind="between", group="SE3", vars=(i, j), aux=M[e], weight=("diag", w)))
    spec["costs"].append(dict(kind="local", group="SE3", vars=(0,), aux=P[0], weight=("scale", np.full((1, 1), 1e-3))))
    return spec


def cpu_baseline_run(data, sample_items, iters=LM_ITERS):
    """Oracle port (numpy restatement of the reference's CPU path: dense A, A^T A by BLAS, dense Cholesky) on a bounded
    sample of the workload, all host threads BLAS can use.  Returns seconds for `iters` LM iterations of the sample."""
    from oracle import nls
    spec = oracle_spec(data, slice(0, sample_items))
    try:  # torchrun exports OMP_NUM_THREADS=1: give BLAS all host cores back for the CPU arm
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=os.cpu_count())
    except Exception:
        import contextlib
        ctx = contextlib.nullcontext()
    with ctx:
        t0 = time.perf_counter()
        out = nls.optimize(spec, method="lm", max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0, sample_trace=False, **LM_KW)
        dt = time.perf_counter() - t0
    return dt, out


C5_RINGS, C5_PER_RING, C5_BATCH = 50, 50, 512


def sparse_c5_leg(th, device, rank, world, pg, timed, steps=2, warmup=1, layout=None, supernodal=False):
    """Config C5 beside the headline: sphere-like pose graph (50 rings x 50 = 2 500 SE3 poses, 4 949 edges: sphere2500's counts),
    batch 512 per GPU (weak scaling: 4096 problems on 8 GPUs), LM (10 iterations, same kwargs) + BaspachoSparseSolver (block-sparse
    Cholesky, batch-lane kernels) on SparseLinearization, device-resident inputs.  Returns a dict for the JSON line."""
    import torch
    from theseus_b200.datasets import build_pose_graph_objective, pose_graph_sphere
    data = pose_graph_sphere(C5_RINGS, C5_PER_RING, C5_BATCH, seed=rank)
    objective, poses = build_pose_graph_objective(th, data, device)
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                                max_iterations=LM_ITERS, step_size=1.0, abs_err_tolerance=0, rel_err_tolerance=0, process_group=pg,
                                linear_solver_kwargs=dict(layout=layout, supernodal_solve=bool(supernodal)))
    layer = th.TheseusLayer(opt)
    inputs = {p.name: data["poses"][i].to(device) for i, p in enumerate(poses)}
    out = {}

    def step():
        with torch.no_grad():
            out["values"], out["info"] = layer.forward(inputs, optimizer_kwargs=LM_KW)

    for _ in range(max(warmup, 1)):
        step()
    ms_step = timed(step, steps) / steps
    info = out["info"]
    lin = opt.linear_solver.linearization
    res = dict(workload="C5: sphere-like SE3 pose graph (2 500 poses, 4 949 edges + 1 prior), batch=512 per GPU, LM(10 it, adaptive+ellipsoidal "
                        "damping) + BaspachoSparseSolver (block-sparse Cholesky) on SparseLinearization, fp64, device-resident inputs",
               value=LM_ITERS * 1e3 / ms_step * world, unit="LM iterations/s (one iteration = one LM step of a 512-problem batch; aggregate over GPUs)",
               ms_per_step=ms_step, steps=steps, warmup=max(warmup, 1), batch_per_gpu=C5_BATCH, global_batch=C5_BATCH * world,
               num_poses=len(poses), num_edges=len(data["edges"]), rows=int(lin.num_rows), cols=int(lin.num_cols),
               problem_iterations_per_s=LM_ITERS * 1e3 / ms_step * world * C5_BATCH, layout=opt.linear_solver.layout_for(C5_BATCH), supernodal_solve=bool(supernodal),
               final_err_mean=float(info.last_err.mean().item()))
    try:
        res["symbolic"] = {k: float(v) for k, v in dict(opt.linear_solver.symbolic_stats).items()}
    except Exception:
        pass
    try:  # phase split of one iteration (same calls as scratch/bench_sparse.py, which produced profiles/r01f_*)
        lam = torch.full((C5_BATCH,), 1e-3, dtype=torch.float64, device=device)
        ms_lin = timed(lin.linearize, 3) / 3
        ms_solve = timed(lambda: opt.linear_solver.solve(damping=lam, ellipsoidal_damping=True, damping_eps=1e-8), 3) / 3
        res["ms_linearize"], res["ms_linear_solve"] = ms_lin, ms_solve
        if "flops" in res.get("symbolic", {}):
            res["factor_gflops_per_item"] = res["symbolic"]["flops"] / 1e9
    except Exception as e:  # the split is a by-product; all ranks take the same path (deterministic), so no rank is left in a barrier
        res["phase_split_error"] = repr(e)[:200]
    return res


FIRST_RUN_LAYOUTS = (("lane_root", "lane_root", False), ("lane_tiled_root", "lane_tiled_root", False),
                     ("lane_tiled_root+supernodal_solve", "lane_tiled_root", True))


def run_c5_layout_child(args):
    """`bench.py --c5-layout L [--c5-supernodal]`: the C5 workload with ONE explicit sparse layout on cuda:0, one JSON line.  Used by the
    parent run (N=1) for the opt-in layouts that had not run on a device when round 1's GPU budget ended -- a separate process, so a
    failing kernel cannot touch the parent's CUDA context or its headline numbers."""
    import torch
    import theseus_b200 as th
    torch.cuda.set_device(0)
    device = torch.device("cuda", 0)

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return float(e0.elapsed_time(e1))

    res = sparse_c5_leg(th, device, 0, 1, None, timed, layout=args.c5_layout, supernodal=args.c5_supernodal)
    for k in ("workload", "unit", "symbolic"):
        res.pop(k, None)
    print(json.dumps(res))


def first_run_layouts(lane_result, timeout_s=150, total_s=330):
    """N=1 only, after everything else is measured: the opt-in sparse layouts, each in its own process (run_c5_layout_child) under a
    timeout (and all of them under `total_s`, so the default run stays within minutes whatever happens in a child).  Same data as the
    parent's `lane` run (seed 0), so `final_err_mean` must agree with it."""
    out = {}
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    t_start = time.perf_counter()
    for tag, layout, supernodal in FIRST_RUN_LAYOUTS:
        left = total_s - (time.perf_counter() - t_start)
        if left < 30:
            out[tag] = dict(error="skipped: the time set aside for the first-run layouts is used up")
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--c5-layout", layout] + (["--c5-supernodal"] if supernodal else [])
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=min(timeout_s, left), env=env, cwd=ROOT)
            lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and lines:
                res = json.loads(lines[-1])
                ref = (lane_result or {}).get("final_err_mean")
                if ref:
                    res["final_err_rel_diff_vs_lane"] = abs(res["final_err_mean"] - ref) / abs(ref)
                out[tag] = res
            else:
                out[tag] = dict(error=f"exit code {r.returncode}", stderr_tail=r.stderr[-400:])
        except subprocess.TimeoutExpired:
            out[tag] = dict(error=f"timeout after {min(timeout_s, left):.0f} s")
        except Exception as e:
            out[tag] = dict(error=repr(e)[:300])
    return out


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; /root/reference is Python and cannot travel to the
    GPU box) on the host cores, same metric/config; each step = a bounded sample (sample_items of the 256 batch items)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from theseus_b200.datasets import pose_graph_synthetic_3d
    sample = 4
    data = pose_graph_synthetic_3d(NUM_POSES, sample, seed=0)
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_baseline_run(data, sample, iters=1)
    times = []
    for _ in range(args.steps):
        dt, _ = cpu_baseline_run(data, sample)
        times.append(dt)
    t_step_sample = float(np.mean(times))
    t_step_full = t_step_sample * BATCH / sample  # the reference's CPU path is linear in the batch (per-item BLAS calls)
    value = LM_ITERS / t_step_full / 1.0
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=t_step_full * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                config=dict(workload=WORKLOAD, note="CPU oracle port of the reference path (dense A, BLAS A^T A, LAPACK potrf); time of a "
                            f"{sample}-item sample scaled linearly to the 256-item batch"),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=f"{sample} of 256 batch items x {LM_ITERS} LM iterations per step"),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip the separately reported config-C5 (block-sparse) workload")
    ap.add_argument("--no-first-run-layouts", action="store_true", help="skip the child runs of the opt-in sparse layouts (N=1 only)")
    ap.add_argument("--c5-layout", default=None, help="child mode: config C5 with this sparse layout on cuda:0, one JSON line")
    ap.add_argument("--c5-supernodal", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.c5_layout is not None:
        return run_c5_layout_child(args)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        print(f"warning: WORLD_SIZE={world} != --gpus {args.gpus}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD

    from theseus_b200 import _lib
    th, data, objective, poses = build_problem(rank, device)
    lib = _lib.load()
    opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.CholeskyDenseSolver, max_iterations=LM_ITERS, step_size=1.0,
                                abs_err_tolerance=0, rel_err_tolerance=0, process_group=pg,
                                cuda_graph=os.environ.get("THB_BENCH_GRAPH", "1") != "0")
    layer = th.TheseusLayer(opt)
    names_pose = [p.name for p in poses]
    # ---- device-resident inputs (for `value`) and pinned host inputs (for `e2e`) ----
    dev_inputs = {p.name: data["poses"][i].to(device) for i, p in enumerate(poses)}
    edge_names = [cf.measurement.name for cf in objective.cost_functions.values() if hasattr(cf, "measurement")]
    host_poses = data["poses"].pin_memory()
    host_meas = data["meas"].pin_memory()
    dev_pose_buf = torch.empty_like(data["poses"], device=device)
    dev_meas_buf = torch.empty_like(data["meas"], device=device)
    host_out = torch.empty_like(data["poses"]).pin_memory()
    host_err = torch.empty(BATCH, dtype=torch.float64).pin_memory()

    def step_resident():
        with torch.no_grad():
            values, info = layer.forward(dev_inputs, optimizer_kwargs=LM_KW)
        return values, info

    def step_e2e():
        # host -> device copy of this step's inputs (initial poses + edge measurements), public API call, device -> host result
        dev_pose_buf.copy_(host_poses, non_blocking=True)
        dev_meas_buf.copy_(host_meas, non_blocking=True)
        inputs = {n: dev_pose_buf[i] for i, n in enumerate(names_pose)}
        inputs.update({n: dev_meas_buf[e] for e, n in enumerate(edge_names)})
        with torch.no_grad():
            values, info = layer.forward(inputs, optimizer_kwargs=LM_KW)
        torch.stack([values[n] for n in names_pose], 0, out=dev_pose_buf)
        host_out.copy_(dev_pose_buf, non_blocking=True)
        host_err.copy_(info.last_err, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return info

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(max(args.warmup, 3)):
        values, info = step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = _lib.total_launches()
    ms_total = timed(step_resident, args.steps)
    launches = int(_lib.total_launches() - l0)  # direct launches + kernels replayed from the captured iteration graph
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    # whole-job aggregate: every rank advances its own 256-problem batch, so the job completes `world` batched LM iterations
    # per iteration time (weak scaling: N=1 value x N is the ideal)
    value = LM_ITERS * 1e3 / ms_step * world
    final_err = info.last_err.mean().item()

    # ---- e2e ----
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps) / args.steps
    h2d = host_poses.numel() * 8 + host_meas.numel() * 8
    d2h = host_out.numel() * 8 + host_err.numel() * 8

    # ---- roofline of the dominant kernel: chol_col_kernel (12 launches = one batched factorisation) ----
    lin = opt.linear_solver.linearization
    lin.linearize()
    AtA = lin.AtA
    B, n = AtA.shape[0], AtA.shape[1]
    alpha = torch.full((B,), 1e-3, dtype=torch.float64, device=device)
    beta = torch.full((B,), 1e-8, dtype=torch.float64, device=device)
    infot = torch.empty(B, dtype=torch.int32, device=device)
    need = int(lib.thb_potrf_workspace_bytes(B, n))
    ws = torch.empty(need, dtype=torch.uint8, device=device)

    def factor():
        _lib.check(lib.thb_potrf_f64(_lib.ptr(AtA), _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(infot), B, n, _lib.ptr(ws), need, _lib.stream_ptr()), "potrf")

    for _ in range(3):
        factor()
    reps = 10
    ms_factor = timed(lambda: factor(), reps) / reps
    nblk = (n + 127) // 128
    flops = B * (n ** 3) / 3.0
    achieved_tf = flops / (ms_factor * 1e-3) / 1e12
    # fp64 peak: MEASURED_PEAKS.json has no fp64 entry -> cuBLAS dgemm 8192^3 measured live, same method as the driver's bf16 peak
    a = torch.randn(8192, 8192, dtype=torch.float64, device=device)
    bm = torch.randn(8192, 8192, dtype=torch.float64, device=device)
    torch.matmul(a, bm)
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, bm)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    peak_tf = 2 * 8192 ** 3 / (best * 1e-3) / 1e12
    del a, bm
    peaks, peaks_src = _measured_peaks()

    # ---- config C5 (block-sparse path), reported beside the headline; every rank runs it (weak scaling) ----
    c5 = None
    if not args.no_c5:
        try:
            c5 = sparse_c5_leg(th, device, rank, world, pg, timed)
        except Exception as e:
            c5 = dict(error=repr(e)[:300])

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N=1 only): oracle port on a bounded sample ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = 4
        dt, out = cpu_baseline_run(data, sample)
        t_full = dt * BATCH / sample
        cpu = dict(value=LM_ITERS / t_full, unit=UNIT, cores=os.cpu_count(), kind="port",
                   sample=f"{sample} of {BATCH} batch items x {LM_ITERS} LM iterations ({dt:.1f} s measured), scaled linearly in batch",
                   final_err_mean_sample=float(out["err_history"][:, -1].mean()))

    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
        config=dict(workload=WORKLOAD, batch_per_gpu=BATCH, global_batch=BATCH * world, num_poses=NUM_POSES,
                    num_edges=len(data["edges"]), rows=int(lin.num_rows), cols=int(lin.num_cols), lm_iterations_per_step=LM_ITERS,
                    problem_iterations_per_s=value * BATCH,
                    l2="working set per iteration (AtA+L = 9.7 GB) >> 126 MB L2, no flush needed",
                    parallelism=f"batch sharded over {world} GPU(s); one all-reduce of 2 int32 per LM iteration",
                    cuda_graph=bool(opt.cuda_graph)),
        clocks=clocks,
        e2e=dict(value=LM_ITERS * 1e3 / ms_e2e * world, unit=UNIT, ms_per_step=ms_e2e, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
        gpu_launches=launches,
        roofline=dict(bound="tensor", kernel="chol_col_kernel (fp64 DMMA left-looking Cholesky, one launch = one batched factorisation of 256 matrices)", achieved=achieved_tf,
                      peak=peak_tf, unit="TFLOP/s", frac=achieved_tf / peak_tf,
                      # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` capture of this kernel at this size
                      # (profiles/r01g_chol_col_ncu_full_details.txt): 31.61 GB + 2.78 GB; algorithmic bytes 4.8 GB -- the left-looking
                      # panels are re-streamed (2.7 TB/s, L2 hit 49 %), the kernel is bound by the FP64 tensor pipe (70.7 % active), not by HBM
                      traffic=34.40e9, traffic_unit="bytes per launch (ncu, B=256, n=1536)", algorithmic_bytes=8.0 * BATCH * lin.num_cols ** 2,
                      flops_per_factorisation=flops, ms_per_factorisation=ms_factor,
                      peak_source="fp64 cuBLAS dgemm 8192^3 measured live in this run (MEASURED_PEAKS.json carries no fp64 figure; "
                                  f"its bf16/HBM entries [{peaks_src}]: {peaks.get('bf16_tflops')} TF/s, {peaks.get('hbm_gbs')} GB/s)",
                      share_of_step=ms_factor * LM_ITERS / ms_step),
        cpu_baseline=cpu, final_err_mean=final_err, sparse_c5=c5)
    if world == 1 and c5 is not None and "error" not in c5 and not args.no_first_run_layouts:
        c5["first_run_layouts"] = first_run_layouts(c5)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
