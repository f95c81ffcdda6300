import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import theseus_b200 as th
from test_gpu_sparse_solver import _random_structure, _dense_system
B, num_cols, sizes, fill, ordering = 1, 40, [1, 2, 3, 6], 0.05, "mindeg"
rng = np.random.default_rng(7 * B + num_cols)
S = _random_structure(rng, num_cols, sizes, fill, num_rows_blocks=3 * num_cols)
A_val = torch.from_numpy(rng.standard_normal((B, S.nnz))).cuda()
b = torch.from_numpy(rng.standard_normal((B, S.num_rows))).cuda()
alpha = torch.from_numpy(rng.random(B) * 0.1).cuda()
xs = {}
for layout in ("lane", "item"):
    solver = th.BaspachoSparseSolver.from_structure(S, ordering=ordering, layout=layout)
    solver.linearization.A_val, solver.linearization.b = A_val, b
    xs[layout] = (solver.solve(damping=alpha, ellipsoidal_damping=True, damping_eps=1e-6).cpu().numpy(),
                  solver.solve(damping=0.37, ellipsoidal_damping=False).cpu().numpy(), solver.solve().cpu().numpy())
AtA, Atb = _dense_system(S, A_val, b)
idx = np.arange(S.num_cols)
for k, (mul, add) in enumerate(((1 + alpha.cpu().numpy()[:, None], 1e-6), (1.0, 0.37), (1.0, 0.0))):
    M = AtA.copy(); M[:, idx, idx] = M[:, idx, idx] * mul + add
    ref = np.linalg.solve(M, Atb[..., None])[..., 0]
    print(k, "cond", np.linalg.cond(M).max(), "lane-ref", np.abs(xs["lane"][k] - ref).max(), "item-ref", np.abs(xs["item"][k] - ref).max(),
          "lane-item", np.abs(xs["lane"][k] - xs["item"][k]).max(), "|x|", np.abs(ref).max())
