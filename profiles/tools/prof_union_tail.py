"""Where does the tail of db_union_dense come from?  Needs an instrumented build of dbscan.cu: profiles/tools/union_profile.patch
(made against commit f9328e0, before the candidate look-ahead; per-thread clocks / scan iterations / find hops / cell visits)."""
import ctypes as C, sys, numpy as np, torch
sys.path.insert(0, '.')
from lidar_ai_recommendation_software_b200 import ops, preprocess as pre, synth, _capi
f = synth.ring_sequence_frame(1)
d = torch.from_numpy(np.ascontiguousarray(f[:, :3], dtype=np.float64)).cuda()
desc, inl, col, ng, idx, X = ops.preprocess_front(d, want_colors=False)
m = ng.shape[0]
lo, hi = np.array(desc.bbox_ng[:3]), np.array(desc.bbox_ng[3:])
for rep in range(2):
    labels, nc, g = ops.dbscan(ng, 0.3, 5, tol=0.0, bounds=(lo, hi))
n = min(m, 1 << 21)
buf = np.zeros(4 * n, dtype=np.int32)
_capi.lib.lidar_debug_union_stats.argtypes = [C.c_void_p, C.c_longlong]
rc = _capi.lib.lidar_debug_union_stats(buf.ctypes.data, 4 * n)
st = buf.reshape(n, 4).astype(np.int64)
clk = st[:, 0] * 64
print('m', m, 'clusters', nc, 'rc', rc)
print('clocks: mean %.0f  p50 %.0f  p99 %.0f  p99.9 %.0f  max %.0f' % (clk.mean(), *np.percentile(clk, [50, 99, 99.9]), clk.max()))
for name, k in (('scan iterations', 1), ('find hops', 2), ('cell visits', 3)):
    v = st[:, k]
    print(f'{name:16s} total {v.sum():>12d}  mean {v.mean():8.1f}  p99 {np.percentile(v, 99):8.0f}  max {v.max():8d}')
top = np.argsort(-clk)[:12]
print('slowest threads: pos clocks scan hops visits')
for t in top:
    print(int(t), int(clk[t]), *[int(x) for x in st[t, 1:]])
# correlation of the time with each count
for name, k in (('scan', 1), ('hops', 2), ('visits', 3)):
    print('corr(clocks,', name, ') =', round(float(np.corrcoef(clk, st[:, k])[0, 1]), 3))
# per-warp maxima (a warp is as slow as its slowest lane)
w = n // 32
wc = clk[:w * 32].reshape(w, 32).max(1)
print('warp clocks: mean %.0f p99 %.0f max %.0f' % (wc.mean(), np.percentile(wc, 99), wc.max()))
