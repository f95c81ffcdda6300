# Library baselines on B200 (vendor path the reference reaches through torch).
import torch, time, json
dev = "cuda"
def ev(f, n=3):
    f(); torch.cuda.synchronize()
    s = torch.cuda.Event(True); e = torch.cuda.Event(True)
    best = 1e30
    for _ in range(n):
        s.record(); f(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best
out = {}
a = torch.randn(8192, 8192, device=dev, dtype=torch.float64); b = torch.randn(8192, 8192, device=dev, dtype=torch.float64)
ms = ev(lambda: a @ b); out["dgemm_8192_tflops"] = 2 * 8192**3 / ms * 1e-9
a4 = torch.randn(4096, 4096, device=dev, dtype=torch.float64)
ms = ev(lambda: a4 @ a4); out["dgemm_4096_tflops"] = 2 * 4096**3 / ms * 1e-9
del a, b, a4
B, n = 64, 1536
M = torch.randn(B, n, n, device=dev, dtype=torch.float64)
S = M @ M.transpose(1, 2) + n * torch.eye(n, device=dev, dtype=torch.float64)
del M
ms = ev(lambda: torch.linalg.cholesky(S)); out["torch_cholesky_B64_n1536_ms"] = ms
out["torch_cholesky_gflops"] = B * n**3 / 3 / ms * 1e-6
L = torch.linalg.cholesky(S); rhs = torch.randn(B, n, 1, device=dev, dtype=torch.float64)
ms = ev(lambda: torch.cholesky_solve(rhs, L)); out["torch_cholesky_solve_B64_ms"] = ms
del S, L
Bb, m = 16, 3120
A = torch.randn(Bb, m, n, device=dev, dtype=torch.float64)
ms = ev(lambda: A.transpose(1, 2).bmm(A)); out["torch_bmm_AtA_B16_ms"] = ms
out["torch_bmm_tflops"] = 2 * Bb * m * n * n / ms * 1e-9
x = torch.empty(1 << 28, device=dev, dtype=torch.float64); y = torch.empty_like(x)
ms = ev(lambda: y.copy_(x)); out["copy_gbs"] = 2 * x.numel() * 8 / ms * 1e-6
ms = ev(lambda: y.zero_()); out["memset_gbs"] = x.numel() * 8 / ms * 1e-6
print(json.dumps(out, indent=1))
