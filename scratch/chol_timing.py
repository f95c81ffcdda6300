import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "prof_chol.py")).read().split("D = AtA.clone()")[0])
raw = C.CDLL(os.environ["THB200_LIB"])
N = 16 * 39936
buf = np.zeros(N, dtype=np.int64)
rc = raw.thb_debug_chol_timing(buf.ctypes.data_as(C.c_void_p), C.c_int64(N)); print("rc", rc)
t = buf.reshape(-1, 16).astype(np.float64)
ntr, nb, Bn = 12, 24, B
starts = np.concatenate([[0], np.cumsum([(ntr - (j >> 1)) * Bn for j in range(nb)])])
clk = 1.965e3  # cycles per us
print("col  nCTA | diag: init kloop B diagC/D total | offdiag: init+dep kloop phaseB flagwait trsm fence total  (us, mean)")
for j in range(nb):
    seg = t[starts[j]:starts[j + 1]]
    d, o = seg[:Bn], seg[Bn:]
    def row(x, diag):
        if len(x) == 0: return ""
        if diag:
            end = np.where(x[:, 6] > 0, x[:, 6], x[:, 4])  # odd columns return after the flag release
            return f"{(x[:,1]-x[:,0]).mean()/clk:5.1f} {(x[:,2]-x[:,1]).mean()/clk:6.1f} {(x[:,3]-x[:,2]).mean()/clk:4.1f} {(np.maximum(x[:,4],x[:,3])-x[:,3]).mean()/clk:5.1f} {(end-x[:,0]).mean()/clk:6.1f}"
        return f"{(x[:,1]-x[:,0]).mean()/clk:5.1f} {(x[:,2]-x[:,1]).mean()/clk:6.1f} {(x[:,3]-x[:,2]).mean()/clk:4.1f} {(x[:,4]-x[:,3]).mean()/clk:5.1f} {(x[:,5]-x[:,4]).mean()/clk:5.1f} {(x[:,6]-x[:,5]).mean()/clk:4.1f} {(x[:,6]-x[:,0]).mean()/clk:6.1f}"
    print(f"{j:3d} {len(seg):5d} | {row(d, True)} | {row(o, False)}")

print("diag internals (us): potrf32#1 | panel+trailing | potrf32#2 | L store + inverse assembly | W store..flag")
for j in (0, 6, 12, 18):
    d = t[starts[j]:starts[j] + Bn]
    print(j, [round(float(((d[:, b2] - d[:, a2]).mean()) / clk), 1) for a2, b2 in ((3, 8), (8, 9), (9, 10), (10, 11), (11, 4))])
