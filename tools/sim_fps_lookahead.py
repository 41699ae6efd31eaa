"""Planning aid: exact K-sample look-ahead for farthest point sampling (how many cluster exchanges a 120k -> 512 sampling\nneeds when the runner-up candidates are retired whenever the argmax leaves them untouched); results are bit-identical to\nthe sequential definition for every K."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from pointcloud_style_transfer_b200 import synthetic as S
def sim(x, npoint, start, K):
    N=x.shape[0]
    dist=np.full(N,1e10,np.float32)
    far=start; out=[far]; rounds=0
    # round: update with 'far' then take top-K of dist; greedily accept
    pending=[far]
    while len(out)<npoint:
        rounds+=1
        for c in pending:
            d=((x-x[c])**2).sum(1).astype(np.float32)
            dist=np.minimum(dist,d)
        # top-K by (value desc, index asc)
        order=np.lexsort((np.arange(N), -dist))[:K]
        acc=[order[0]]
        for j in range(1,K):
            c=order[j]
            ok = dist[c]>0
            for a in acc:
                dd=np.float32(((x[c]-x[a])**2).sum())
                if not (dd>=dist[c]): ok=False;break
            if not ok: break
            acc.append(c)
        acc=acc[:npoint-len(out)]
        out+=list(acc); pending=acc
    return np.array(out), rounds
def ref(x,npoint,start):
    N=x.shape[0]; dist=np.full(N,1e10,np.float32); far=start; out=[]
    for i in range(npoint):
        out.append(far)
        d=((x-x[far])**2).sum(1).astype(np.float32); dist=np.minimum(dist,d); far=int(np.argmax(dist))
    return np.array(out)
for name,x in (("lidar",S.lidar_scan(0).numpy()[0]),("uniform",S.uniform_cloud(0,1,120000).numpy()[0]), ("lidar16k", S.lidar_scan(1,16384).numpy()[0])):
    r=ref(x,512,1234)
    for K in (2,3,4,8):
        o,rounds=sim(x,512,1234,K)
        print(name,"K",K,"rounds",rounds,"exact",np.array_equal(o,r))
