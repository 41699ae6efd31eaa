"""CPU simulation (numpy, 1500 sampled queries) of an EXACT grid-based 3-NN on the synthetic LiDAR scan: references binned in a
uniform grid of cell size h, cells visited in Chebyshev rings until the k-th distance is covered by the visited volume.
Prints evaluated pairs and probed cells per query -- the brute-force sweep evaluates all 30 000 references per query.
Planning aid for the next kNN kernel; nothing in the product path uses it."""
import numpy as np, torch, sys, time
sys.path.insert(0,'/root/repo')
from pointcloud_style_transfer_b200 import synthetic as S
x = S.lidar_scan(0)[0].numpy().astype(np.float64)   # [120000,3]
rng = np.random.default_rng(0)
perm = rng.permutation(len(x))
ref = x[perm[:30000]]; qry = x[perm[30000:]]
k = 3
for _ in (0,):
    lo, hi = ref.min(0), ref.max(0)
    for h in (0.005, 0.01, 0.02, 0.05):
        ci = np.floor((ref - lo) / h).astype(np.int64)
        dims = ci.max(0) + 1
        key = (ci[:,0]*dims[1] + ci[:,1])*dims[2] + ci[:,2]
        order = np.argsort(key, kind='stable'); keys = key[order]
        uniq, start, cnt = np.unique(keys, return_index=True, return_counts=True)
        cellmap = dict(zip(uniq.tolist(), zip(start.tolist(), cnt.tolist())))
        rs = ref[order]
        sample = qry[rng.choice(len(qry), int(__import__("os").environ.get("QUERIES", "1500")), replace=False)]
        pairs = []; rings = []; cells_visited = []
        for q in sample:
            qc = np.floor((q - lo) / h).astype(np.int64)
            frac = (q - lo) / h - qc
            margin = min(frac.min(), (1 - frac).min()) * h
            best = np.full(k, np.inf); npairs = 0; ncell = 0
            r = 0
            while True:
                # ring r (Chebyshev distance == r)
                rngs = range(-r, r + 1)
                for dx in rngs:
                    for dy in rngs:
                        for dz in rngs:
                            if max(abs(dx), abs(dy), abs(dz)) != r: continue
                            c = qc + (dx, dy, dz)
                            if (c < 0).any() or (c >= dims).any(): continue
                            ncell += 1
                            e = cellmap.get(int((c[0]*dims[1] + c[1])*dims[2] + c[2]))
                            if e is None: continue
                            pts = rs[e[0]:e[0]+e[1]]
                            d = np.sqrt(((pts - q) ** 2).sum(1)); npairs += len(d)
                            best = np.sort(np.concatenate([best, d]))[:k]
                if best[k-1] <= r * h + margin or r > 60: break
                r += 1
            pairs.append(npairs); rings.append(r); cells_visited.append(ncell)
        print(f"h={h}: occupied cells {len(uniq)}, mean pts/cell {cnt.mean():.1f}; per query: pairs mean {np.mean(pairs):.0f} p99 {np.percentile(pairs,99):.0f} max {max(pairs)}; rings mean {np.mean(rings):.1f} max {max(rings)}; cells probed mean {np.mean(cells_visited):.0f} max {max(cells_visited)}")
