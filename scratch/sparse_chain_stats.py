import sys, numpy as np
sys.path.insert(0,'/root/repo')
from theseus_b200.datasets import pose_graph_sphere, pose_graph_synthetic_3d
from theseus_b200.sparse import analyze
def plan_of(data):
    N=len(data["poses"]); adj=[set([i]) for i in range(N)]
    for e in data["edges"]:
        i,j=int(e[0]),int(e[1]); adj[i].add(j); adj[j].add(i)
    ptrs=[0]; inds=[]
    for i in range(N):
        inds+=sorted(adj[i]); ptrs.append(len(inds))
    return analyze(np.full(N,6),np.array(ptrs),np.array(inds))
for name,data in (("c5",pose_graph_sphere(50,50,1)),("c2",pose_graph_synthetic_3d(256,1))):
    P=plan_of(data); N=P.N
    st=[set(int(x) for x in s) for s in P.struct]
    parent=[int(P.struct[j][0]) if len(P.struct[j]) else -1 for j in range(N)]
    for maxw in (1,2,4,8,64):
        chain_of=[-1]*N; chains=[]
        for j in range(N):
            if chain_of[j]>=0: continue
            c=[j]; chain_of[j]=len(chains)
            while len(c)<maxw:
                k=c[-1]; p=parent[k]
                if p!=k+1 or st[k]!=({p}|st[p]): break
                # p must have k as its only child to keep the chain fundamental
                if sum(1 for q in range(N) if parent[q]==p)!=1: break
                c.append(p); chain_of[p]=len(chains)
            chains.append(c)
        # chain levels
        nc=len(chains); lev=[0]*nc
        for ci,c in enumerate(chains):
            p=parent[c[-1]]
            if p>=0:
                pc=chain_of[p]; lev[pc]=max(lev[pc],lev[ci]+1)
        flops=[0]*nc
        print(name,"maxw",maxw,"chains",nc,"levels",max(lev)+1,"widths>1:",sum(1 for c in chains if len(c)>1),"max width",max(len(c) for c in chains))
