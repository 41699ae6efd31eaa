"""CPU simulation (numpy, 800 sampled queries) of an EXACT grid-based nearest-neighbour search for the 120k x 120k Chamfer sweep
(scan 0 against scan 100): the target binned in a uniform grid of cell size h, cells visited in Chebyshev rings until the best squared
distance (+ 2e-6 for the fp32 rounding of the reference formula) is covered.  Prints evaluated pairs and probed cells per query; the
brute-force sweep evaluates 120 000 pairs per query.  Planning aid; nothing in the product path uses it."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from pointcloud_style_transfer_b200 import synthetic as S
a = S.lidar_scan(0)[0].numpy().astype(np.float64); b = S.lidar_scan(100)[0].numpy().astype(np.float64)
rng = np.random.default_rng(0)
for h in (0.004, 0.008, 0.016):
    lo = np.minimum(a.min(0), b.min(0))
    ci = np.floor((b - lo) / h).astype(np.int64); dims = ci.max(0) + 2
    key = (ci[:,0]*dims[1] + ci[:,1])*dims[2] + ci[:,2]
    order = np.argsort(key, kind='stable'); keys = key[order]
    uniq, start, cnt = np.unique(keys, return_index=True, return_counts=True)
    cellmap = dict(zip(uniq.tolist(), zip(start.tolist(), cnt.tolist()))); rs = b[order]
    pairs=[]; cells=[]; rings=[]
    for q in a[rng.choice(len(a), 800, replace=False)]:
        qc = np.floor((q - lo) / h).astype(np.int64); frac = (q - lo)/h - qc
        margin = min(frac.min(), (1-frac).min())*h
        best = np.inf; npairs=0; ncell=0; r=0
        while True:
            R = range(-r, r+1)
            for dx in R:
                for dy in R:
                    for dz in R:
                        if max(abs(dx),abs(dy),abs(dz)) != r: continue
                        c = qc + (dx,dy,dz)
                        if (c<0).any() or (c>=dims).any(): continue
                        ncell += 1
                        e = cellmap.get(int((c[0]*dims[1]+c[1])*dims[2]+c[2]))
                        if e is None: continue
                        d2 = ((rs[e[0]:e[0]+e[1]] - q)**2).sum(1); npairs += len(d2); best = min(best, d2.min())
            if best + 2e-6 <= (r*h + margin)**2 or r > 80: break
            r += 1
        pairs.append(npairs); cells.append(ncell); rings.append(r)
    print(f"h={h}: occupied {len(uniq)} cells, {cnt.mean():.1f} pts/cell; per query pairs mean {np.mean(pairs):.0f} p99 {np.percentile(pairs,99):.0f} max {max(pairs)}; cells mean {np.mean(cells):.0f} p99 {np.percentile(cells,99):.0f} max {max(cells)}; rings mean {np.mean(rings):.1f} max {max(rings)}")
