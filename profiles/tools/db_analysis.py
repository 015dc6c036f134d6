import sys, time, numpy as np
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import synth
from scipy.spatial import cKDTree
f = synth.ring_sequence_frame(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
p = f[:, :3].astype(np.float64)
m0 = (np.abs(p - p.mean(0)) < 3 * p.std(0)).all(1); p = p[m0]
thr = np.percentile(p[:, 2], 30); ng = p[p[:, 2] > thr]
print('non-ground', len(ng))
eps = 0.3; c = eps / np.sqrt(3) * (1 - 1e-6)
mn = ng.min(0)
ijk = np.floor((ng - mn) / c).astype(np.int64)
g = ijk.max(0) + 1
cid = (ijk[:, 0] * g[1] + ijk[:, 1]) * g[2] + ijk[:, 2]
order = np.argsort(cid, kind='stable'); sc = cid[order]
u, start, cnt = np.unique(sc, return_index=True, return_counts=True)
print('cells', g, 'occupied', len(u), 'count pct', np.percentile(cnt, [50, 90, 99, 99.9]), 'max', cnt.max())
print('points in cells >=64:', cnt[cnt >= 64].sum() / len(ng), ' cells>=64:', (cnt >= 64).sum(), ' cells >=5 (full):', (cnt >= 5).sum())
# heavy cells: forward neighbour pairs among heavy cells
heavy = u[cnt >= 64]; hset = {int(x): i for i, x in enumerate(heavy)}
pts_of = lambda cell: ng[order[start[np.searchsorted(u, cell)]: start[np.searchsorted(u, cell)] + cnt[np.searchsorted(u, cell)]]]
pairs = 0; nonmerge = 0; w_all = 0; w_non = 0; big = []
t0 = time.time()
offs = [(a, b, d) for a in range(-2, 3) for b in range(-2, 3) for d in range(-2, 3)]
for cell in heavy:
    cz = cell % g[2]; t = cell // g[2]; cy = t % g[1]; cx = t // g[1]
    A = pts_of(cell); ta = None
    for (a, b, d) in offs:
        nx, ny, nz = cx + a, cy + b, cz + d
        if nx < 0 or ny < 0 or nz < 0 or nx >= g[0] or ny >= g[1] or nz >= g[2]: continue
        nc = (nx * g[1] + ny) * g[2] + nz
        if nc <= cell or int(nc) not in hset: continue
        B = pts_of(nc)
        if ta is None: ta = cKDTree(A)
        dmin = cKDTree(B).query(A, k=1)[0].min()
        pairs += 1; w = len(A) * len(B); w_all += w
        if dmin > eps:
            nonmerge += 1; w_non += w; big.append((w, len(A), len(B), (a, b, d), round(float(dmin), 3)))
print('heavy-heavy forward pairs', pairs, 'non-merging', nonmerge, 'weight all %.3g non %.3g' % (w_all, w_non), 'time', time.time() - t0)
big.sort(reverse=True); print(big[:15])
