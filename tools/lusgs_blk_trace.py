"""Development aid: per-tile timeline of one k_lusgs_blk launch (ICSB200_LUSGS_PROF=1 ICSB200_LUSGS_TRACE=1 are set here).
usage: lusgs_blk_trace.py n | lusgs_blk_trace.py bump nxb ny.  Prints, for the forward sweep, where the time between a tile's
dependencies being published and the tile being published goes, and how far ahead of its consumers the halo warp runs."""
import ctypes as C
import os, sys
import numpy as np
os.environ["ICSB200_LUSGS_PROF"] = "1"
os.environ["ICSB200_LUSGS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import cases
from icsfoam_b200.context import Context, lib
if len(sys.argv) > 1 and sys.argv[1] == "bump":
    case = cases.bump(int(sys.argv[2]), int(sys.argv[3]))
else:
    case = cases.onera_box(int(sys.argv[1]) if len(sys.argv) > 1 else 128)
g = case.apply(Context())
info = g.schedule_info()
print("cells", case.mesh.n_cells, "schedule", info)
g.calc_flux(); g.residual(); g.pseudo_dt(); g.assemble()
N = case.mesh.n_cells
x = (np.ones(N), np.ones((N, 3)), np.ones(N))
for _ in range(3):
    g.precondition("LUSGS", *x)
nT = info["n_tiles"]
L = lib()
tr = np.zeros((2 * nT, 8), np.int64)
L.icsb200_debug_blk_trace.restype = C.c_int
n = L.icsb200_debug_blk_trace(g.h, tr.ctypes.data_as(C.c_void_p), 2 * nT)
deps = np.zeros((nT, 8), np.int32)
L.icsb200_debug_blk_deps.restype = C.c_int
L.icsb200_debug_blk_deps(g.h, deps.ctypes.data_as(C.c_void_p), nT)
f = tr[:nT].astype(float)
t0 = f[:, 0].min()
meta, seen, halo, start, fin, pub, cta = (f[:, k] - t0 for k in (0, 1, 2, 3, 4, 5, 7))
cta = f[:, 7].astype(int)
print("forward sweep: %.3f ms, reverse sweep ends %.3f ms after the first forward tile" % ((pub.max()) / 1e6, (tr[nT:, 5].max() - t0) / 1e6))
lastdep = np.full(nT, np.nan)
for t in range(nT):
    d = deps[t][deps[t] >= 0]
    d = d[d < nT]
    if len(d): lastdep[t] = pub[d].max()
ok = ~np.isnan(lastdep)
q = lambda a: "median %.2f  p10 %.2f  p90 %.2f us" % (np.median(a) / 1e3, np.percentile(a, 10) / 1e3, np.percentile(a, 90) / 1e3)
print("sweep (consumers start -> finish):        ", q(fin - start))
print("publish (finish -> flag):                 ", q(pub - fin))
print("last dependency published -> flags seen:  ", q((seen - lastdep)[ok]))
print("halo warp: metadata -> flags seen:        ", q(seen - meta))
print("halo warp: flags seen -> halo staged:     ", q(halo - seen))
print("halo staged -> consumers start (slack):   ", q(start - halo))
print("last dependency published -> own publish: ", q((pub - lastdep)[ok]))
# per CTA: gap between consecutive tiles' sweeps
order = np.lexsort((start, cta))
gaps = []
for a, b in zip(order[:-1], order[1:]):
    if cta[a] == cta[b]: gaps.append(start[b] - fin[a])
print("idle between consecutive sweeps of a CTA: ", q(np.array(gaps)))
dist = []
for t in range(0, nT, max(1, nT // 2000)):
    d = deps[t][deps[t] >= 0]
    if len(d): dist.append(t - d.max())
print("tile index distance to the nearest dependency: median %d, p10 %d, min %d" % (np.median(dist), np.percentile(dist, 10), np.min(dist)))
if os.environ.get("ICS_TRACE_SAVE"):
    np.savez_compressed(os.environ["ICS_TRACE_SAVE"], trace=tr, deps=deps)
