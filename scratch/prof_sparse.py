# One linear solve of the block-sparse solver inside a cudaProfilerStart/Stop range (for ncu --profile-from-start off).
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import theseus_b200 as th
from theseus_b200.datasets import pose_graph_synthetic_3d, pose_graph_sphere, build_pose_graph_objective
which, B, layout = sys.argv[1], int(sys.argv[2]), sys.argv[3]
data = pose_graph_sphere(50, 50, B) if which == "c5" else pose_graph_synthetic_3d(256, B)
objective, poses = build_pose_graph_objective(th, data, torch.device("cuda", 0))
opt = th.LevenbergMarquardt(objective, linear_solver_cls=th.BaspachoSparseSolver, linearization_cls=th.SparseLinearization,
                            max_iterations=2, abs_err_tolerance=0, rel_err_tolerance=0, linear_solver_kwargs=dict(layout=layout))
objective.update({p.name: data["poses"][i].cuda() for i, p in enumerate(poses)})
lin = opt.linear_solver.linearization
lin.linearize()
lam = torch.full((B,), 1e-3, dtype=torch.float64, device="cuda")
opt.linear_solver.solve(damping=lam, ellipsoidal_damping=True, damping_eps=1e-8)
torch.cuda.synchronize()
torch.cuda.profiler.start()
opt.linear_solver.solve(damping=lam, ellipsoidal_damping=True, damping_eps=1e-8)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
