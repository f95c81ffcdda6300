import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import nls
from helpers import load, pgo_spec, lm_kwargs_of
g=load("pgo_c5_lm"); method, iters, kw = lm_kwargs_of(g)
t0=time.time()
out=nls.optimize(pgo_spec(g), method=method, max_iterations=iters, abs_err_tolerance=0, rel_err_tolerance=0, **kw)
print("oracle err history", out["err_history"], "ref", g["trace_err"].ravel(), "took", time.time()-t0)
print("rel diff", np.abs(out["err_history"][:,1:].ravel()-g["trace_err"].ravel())/g["trace_err"].ravel())
