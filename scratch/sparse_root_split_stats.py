import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from theseus_b200.datasets import pose_graph_sphere
from theseus_b200.sparse import analyze, root_split
data = pose_graph_sphere(50, 50, 1)
N = len(data["poses"]); adj=[set([i]) for i in range(N)]
for e in data["edges"]:
    i,j=int(e[0]),int(e[1]); adj[i].add(j); adj[j].add(i)
ptrs=[0]; inds=[]
for i in range(N):
    inds+=sorted(adj[i]); ptrs.append(len(inds))
t0=time.time(); P=analyze(np.full(N,6),np.array(ptrs),np.array(inds)); t1=time.time()
sp=root_split(P); t2=time.time()
print("analyze %.2fs root_split %.2fs"%(t1-t0,t2-t1), "cut",sp["cut"],"root cols",N-sp["cut"],"root dof",sp["root_dof"], "bottom launches",len(sp["bottom"]["launches"]),"(all:",len(P.lane["launches"]),")",
      "root assembly pairs", int((sp["ru_p1"]-sp["ru_p0"]).sum()), "of", int(P.stats["num_updates"]), "bottom levels", int(P.level[sp["cut"]]), "of", int(P.stats["levels"]))
for md in (384, 768, 1536):
    s2=root_split(P, max_root_dof=md); print(md, "->", None if s2 is None else (N-s2["cut"], s2["root_dof"]))
