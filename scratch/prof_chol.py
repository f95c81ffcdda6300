# Drives only the batched dense Cholesky (for ncu): B=256 SPD matrices with the C2 size n=1536.
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from theseus_b200 import _lib
lib = _lib.load()
B, n = int(sys.argv[1]) if len(sys.argv) > 1 else 256, int(sys.argv[2]) if len(sys.argv) > 2 else 1536
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
g = torch.Generator(device="cuda").manual_seed(0)
AtA = torch.empty(B, n, n, dtype=torch.float64, device="cuda")
for b0 in range(0, B, 32):
    M = torch.randn(min(32, B - b0), n, n // 4, dtype=torch.float64, device="cuda", generator=g)
    AtA[b0:b0 + M.shape[0]] = M @ M.transpose(1, 2)
    del M
AtA += n * torch.eye(n, dtype=torch.float64, device="cuda")
alpha = torch.full((B,), 1e-3, dtype=torch.float64, device="cuda"); beta = torch.full((B,), 1e-8, dtype=torch.float64, device="cuda")
info = torch.empty(B, dtype=torch.int32, device="cuda")
need = int(lib.thb_potrf_workspace_bytes(B, n)); ws = torch.empty(need, dtype=torch.uint8, device="cuda")
rhs = torch.randn(B, n, dtype=torch.float64, device="cuda", generator=g); x = torch.empty_like(rhs)
torch.cuda.synchronize()
for r in range(reps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    _lib.check(lib.thb_potrf_f64(_lib.ptr(AtA), _lib.ptr(alpha), _lib.ptr(beta), _lib.ptr(info), B, n, _lib.ptr(ws), need, _lib.stream_ptr()), "potrf")
    e1.record()
    _lib.check(lib.thb_potrs_f64(_lib.ptr(rhs), _lib.ptr(x), B, n, _lib.ptr(ws), need, _lib.stream_ptr()), "potrs")
    e2.record()
    torch.cuda.synchronize()
    print(f"rep {r}: potrf {e0.elapsed_time(e1):.3f} ms ({B*n**3/3/e0.elapsed_time(e1)*1e-9:.2f} TF/s)  potrs {e1.elapsed_time(e2):.3f} ms  info_nonzero={int((info!=0).sum())}")
D = AtA.clone(); idx = torch.arange(n, device="cuda"); D[:, idx, idx] += 1e-3 * AtA[:, idx, idx] + 1e-8
res = (torch.bmm(D[:8], x[:8].unsqueeze(2)).squeeze(2) - rhs[:8]).abs().max().item()
print("residual", res)
