"""Development aid: where does the LU-SGS forward sweep wait?  Run with ICSB200_LUSGS_TRACE=1 on the GPU box."""
import ctypes as C, sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import cases
from icsfoam_b200.context import Context, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
case = cases.onera_box(n)
g = case.apply(Context())
g.calc_flux(); g.residual(); g.pseudo_dt(); g.assemble()
N = case.mesh.n_cells
x = (np.ones(N), np.ones((N, 3)), np.ones(N))
for _ in range(3):
    g.precondition("LUSGS", *x)
nS = g.schedule_info()["n_positions"] // 32
tr = np.zeros((nS, 4), np.int64); nl = np.zeros(nS, np.int32); c3 = np.zeros((nS, 3), np.int32)
f = lib().icsb200_debug_lusgs_trace
f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
assert f(g.h, tr.ctypes.data, nl.ctypes.data, c3.ctypes.data) == 0
t0 = tr[:, 0].min()
tr = (tr - t0) / 1e3  # us
print("slices", nS, "forward sweep span us", tr[:, 3].max())
dep = c3 // 32
valid = nl > 0
depstore = np.where(dep >= 0, tr[np.maximum(dep, 0), 3], -1).max(1)
s = np.where(valid)[0]
print("prefetch issue (t1-t0)      median %.2f p90 %.2f" % tuple(np.percentile(tr[s, 1] - tr[s, 0], [50, 90])))
print("wait for deps (t2-t1)       median %.2f p90 %.2f" % tuple(np.percentile(tr[s, 2] - tr[s, 1], [50, 90])))
print("ready after last dep store  median %.2f p90 %.2f p99 %.2f" % tuple(np.percentile(tr[s, 2] - depstore[s], [50, 90, 99])))
print("compute+store (t3-t2)       median %.2f p90 %.2f p99 %.2f" % tuple(np.percentile(tr[s, 3] - tr[s, 2], [50, 90, 99])))
print("start after last dep store  median %.2f (negative = warp was already waiting)" % np.median(tr[s, 0] - depstore[s]))
late = (tr[s, 0] > depstore[s]).mean()
print("fraction of slices whose warp started AFTER deps were stored: %.3f" % late)

# ---- composition of the critical path: walk back from the slice that finished last through its latest dependency
cur = int(np.argmax(tr[:, 3]))
hops, det, comp, latecnt, latewait = 0, [], [], 0, []
while True:
    d = dep[cur][dep[cur] >= 0] if nl[cur] > 0 else []
    if len(d) == 0:
        break
    prev = int(d[np.argmax(tr[d, 3])])
    ds = tr[prev, 3]
    det.append(tr[cur, 2] - ds)
    comp.append(tr[cur, 3] - tr[cur, 2])
    if tr[cur, 0] > ds:
        latecnt += 1
        latewait.append(tr[cur, 2] - tr[cur, 0])
    hops += 1
    cur = prev
det, comp = np.array(det), np.array(comp)
print("critical path: hops %d, span %.1f us; detection mean %.2f (median %.2f p90 %.2f), compute+store mean %.2f (median %.2f p90 %.2f)" % (
    hops, det.sum() + comp.sum(), det.mean(), np.median(det), np.percentile(det, 90), comp.mean(), np.median(comp), np.percentile(comp, 90)))
print("critical path: slices whose warp started after the dependency was stored: %d (their start->ready mean %.2f us)" % (
    latecnt, np.mean(latewait) if latewait else 0.0))
big = np.argsort(det)[-5:]
print("largest detection delays on the critical path (us):", np.round(det[big], 2))
