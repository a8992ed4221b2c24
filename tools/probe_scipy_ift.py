"""Probe of scipy.ndimage.watershed_ift: which sequential model explains its output?
Result (SciPy 1.18.1): bucket-queue IFT with cost max(cost(v), |I(p)-I(v)|) and a FLAT
neighbourhood (flat index + offset in [0, N)), i.e. rows/planes wrap at the volume faces."""
import numpy as np
from scipy import ndimage
from scipy.ndimage import generate_binary_structure

def ift_seq(img, mk, st, costfn):
    """Sequential bucket-queue IFT as remembered; costfn(vcost, Iv, Ip, Iroot) -> candidate cost."""
    img=img.astype(np.int64); shp=img.shape; n=img.size
    flat=img.ravel(); lab=mk.ravel().astype(np.int64).copy()
    maxval=int(flat.max())
    INF=maxval+1
    cost=np.full(n,INF,np.int64); done=np.zeros(n,bool); root=np.arange(n)
    buckets=[[] for _ in range(maxval+2)]   # list used as deque: front = index 0
    for j in range(n):
        if lab[j]!=0:
            cost[j]=0
            if lab[j]>0: buckets[0].insert(0,j)
            else: buckets[0].append(j)
    c=[s//2 for s in st.shape]
    offs=[tuple(np.array(i)-c) for i in np.argwhere(st)]
    for b in range(maxval+1):
        while buckets[b]:
            v=buckets[b].pop(0); done[v]=True
            vz,vy,vx=np.unravel_index(v,shp)
            for o in offs:
                q=(vz+o[0],vy+o[1],vx+o[2])
                if not all(0<=q[i]<shp[i] for i in range(3)): continue
                p=np.ravel_multi_index(q,shp)
                if done[p]: continue
                cand=costfn(cost[v],flat[v],flat[p],flat[root[v]])
                if cand<cost[p]:
                    if cost[p]<=maxval and p in buckets[cost[p]]: buckets[cost[p]].remove(p)
                    cost[p]=cand; lab[p]=lab[v]; root[p]=root[v]
                    if lab[v]<0: buckets[cand].append(p)
                    else: buckets[cand].insert(0,p)
    return lab.reshape(shp)

H={"H1 max(vcost,|Ip-Iv|)": lambda vc,iv,ip,ir: max(vc,abs(ip-iv)),
   "H3 max(vcost,|Ip-Iroot|)": lambda vc,iv,ip,ir: max(vc,abs(ip-ir)),
   "H4 |Ip-Iv|": lambda vc,iv,ip,ir: abs(ip-iv),
   "H5 max(vcost,Ip)": lambda vc,iv,ip,ir: max(vc,ip),
   "H6 vcost+|Ip-Iv|": lambda vc,iv,ip,ir: min(vc+abs(ip-iv), 10**9)}
rng=np.random.default_rng(7)
res={k:[] for k in H}
for t in range(60):
    shape=(2,4,5); img=rng.integers(0,40,shape).astype(np.uint16)
    mk=np.zeros(shape,np.int16); mk.flat[0]=1; mk.flat[-1]=2; mk[1,2,2]=3
    st=generate_binary_structure(3,1 if t%2 else 3)
    w=ndimage.watershed_ift(img,mk,st)
    for k,f in H.items():
        try:
            a=ift_seq(img,mk,st,f); res[k].append((a==w).mean())
        except Exception as e:
            res[k].append(-1)
for k,v in res.items(): print(k, np.round(np.mean(v),3), "exact cases:", sum(x==1.0 for x in v),"/",len(v))

def ift_seq_wrap(img, mk, st, mode):
    img=img.astype(np.int64); shp=img.shape; n=img.size
    flat=img.ravel(); lab=mk.ravel().astype(np.int64).copy()
    maxval=int(flat.max()); INF=maxval+1
    cost=np.full(n,INF,np.int64); done=np.zeros(n,bool)
    buckets=[[] for _ in range(maxval+2)]
    for j in range(n):
        if lab[j]!=0:
            cost[j]=0
            if lab[j]>0: buckets[0].insert(0,j)
            else: buckets[0].append(j)
    c=[s//2 for s in st.shape]
    strides=[shp[1]*shp[2],shp[2],1]
    offs=[tuple(np.array(i)-c) for i in np.argwhere(st)]
    for b in range(maxval+1):
        while buckets[b]:
            v=buckets[b].pop(0); done[v]=True
            vc=np.unravel_index(v,shp)
            for o in offs:
                p=v+o[0]*strides[0]+o[1]*strides[1]+o[2]*strides[2]
                if mode=="flat":
                    if not (0<=p<n): continue
                elif mode=="coord1":   # accept if every coordinate of p differs by <=1 from v
                    if not (0<=p<n): continue
                    pc=np.unravel_index(p,shp)
                    if any(abs(int(pc[i])-int(vc[i]))>1 for i in range(3)): continue
                if done[p]: continue
                cand=max(cost[v],abs(flat[p]-flat[v]))
                if cand<cost[p]:
                    if cost[p]<=maxval and p in buckets[cost[p]]: buckets[cost[p]].remove(p)
                    cost[p]=cand; lab[p]=lab[v]
                    buckets[cand].insert(0,p)
    return lab.reshape(shp)

rng=np.random.default_rng(7)
r2={"flat":[], "coord1":[]}
for t in range(60):
    shape=(2,4,5); img=rng.integers(0,40,shape).astype(np.uint16)
    mk=np.zeros(shape,np.int16); mk.flat[0]=1; mk.flat[-1]=2; mk[1,2,2]=3
    st=generate_binary_structure(3,1 if t%2 else 3)
    w=ndimage.watershed_ift(img,mk,st)
    for m in r2: r2[m].append((ift_seq_wrap(img,mk,st,m)==w).mean())
for k,v in r2.items(): print(k, np.round(np.mean(v),3), "exact:", sum(x==1.0 for x in v),"/",len(v))
