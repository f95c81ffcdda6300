import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from theseus_b200.datasets import pose_graph_sphere
from theseus_b200.sparse import analyze
data = pose_graph_sphere(50, 50, 1)
N = len(data["poses"]); edges = data["edges"]
adj=[set([i]) for i in range(N)]
for e in edges:
    i,j=int(e[0]),int(e[1]); adj[i].add(j); adj[j].add(i)
ptrs=[0]; inds=[]
for i in range(N):
    inds+=sorted(adj[i]); ptrs.append(len(inds))
P=analyze(np.full(N,6),np.array(ptrs),np.array(inds))
struct=[set(int(x) for x in s) for s in P.struct]
# rowlist[j] = set of k with L_jk != 0
rowlist=[set() for _ in range(N)]
for k in range(N):
    for i in struct[k]: rowlist[i].add(k)
pairs_total=0
for TC,TR in ((1,1),(1,4),(2,2),(2,4),(4,4),(4,8)):
    loads=0; updates=0
    # target tiles: column groups of TC consecutive columns; within a column group, rows = union of struct + the columns themselves, grouped by TR consecutive (sorted) rows
    for j0 in range(0,N,TC):
        cols=list(range(j0,min(N,j0+TC)))
        rows=sorted(set().union(*[struct[j]|{j} for j in cols]))
        for r0 in range(0,len(rows),TR):
            rt=rows[r0:r0+TR]
            # targets in tile: (i,j) with i in rt, j in cols, i>=j, block exists (i==j or i in struct[j])
            tg=[(i,j) for j in cols for i in rt if i>=j and (i==j or i in struct[j])]
            if not tg: continue
            ks=set()
            for (i,j) in tg:
                ks |= (rowlist[i]|set()) & rowlist[j] if i!=j else rowlist[j]
            for k in ks:
                ri={i for (i,j) in tg if k in rowlist[i] and k in rowlist[j]}
                cj={j for (i,j) in tg if k in rowlist[i] and k in rowlist[j]}
                nu=sum(1 for (i,j) in tg if k in rowlist[i] and k in rowlist[j])
                loads+=len(ri|cj)  # a block L_xk loaded once per tile even if used as row and col
                updates+=nu
    print(f"TC={TC} TR={TR}: updates {updates}, block loads {loads}, loads per update {loads/updates:.3f} (now 2.0 for off-diag)")
