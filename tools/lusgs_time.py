"""Development aid: schedule statistics and LU-SGS / SpMV kernel times on the GPU box.
usage: lusgs_time.py n | lusgs_time.py bump nxb ny ; ICSB200_LUSGS_PROF=1 adds the per-phase cycle profile of k_lusgs_blk."""
import ctypes as C
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import cases
from icsfoam_b200.context import Context, lib
if len(sys.argv) > 1 and sys.argv[1] == "bump":
    case = cases.bump(int(sys.argv[2]), int(sys.argv[3]))
else:
    case = cases.onera_box(int(sys.argv[1]) if len(sys.argv) > 1 else 128)
g = case.apply(Context())
print("cells", case.mesh.n_cells, "schedule", g.schedule_info())
g.calc_flux(); g.residual(); g.pseudo_dt(); g.assemble()
N = case.mesh.n_cells
x = (np.ones(N), np.ones((N, 3)), np.ones(N))
for _ in range(2):
    g.precondition("LUSGS", *x)
g.timers_reset(True)
for _ in range(5):
    g.precondition("LUSGS", *x)
    g.matrix_mul(*x)
t = g.timers_get()
print({k: (round(v[0] / max(v[1], 1), 3), v[1]) for k, v in t.items() if v[1]})
if os.environ.get("ICSB200_LUSGS_PROF") and g.schedule_info().get("blk"):
    f = lib().icsb200_debug_lusgs_prof
    f.restype = C.c_int
    buf = np.zeros((148, 24), np.int64)
    n = f(g.h, buf.ctypes.data_as(C.c_void_p), 148)
    p = buf[:n].astype(float)
    names = ["meta", "halo", "levels", "lvlwait", "fullwait"]
    tot = p[:, 7].mean()
    print("prof: CTAs %d, mean cycles/CTA %.0f; share of phases (lvlwait, fullwait are parts of levels):" % (n, tot),
          {k: round(p[:, i].mean() / tot, 3) for i, k in enumerate(names)})
    tiles, levs = p[:, 5].sum(), p[:, 6].sum()
    print("prof: cycles per tile-sweep %.0f (meta %.0f halo %.0f levels %.0f lvlwait %.0f fullwait %.0f); cycles per level %.0f" % (
        p[:, 7].sum() / tiles, *(p[:, i].sum() / tiles for i in range(5)), p[:, 2].sum() / levs))
    print("prof: helper warps, cycles per tile-sweep: halo warp [meta wait %.0f, flag polls %.0f, fence %.0f, gather %.0f]; publish warp [wait %.0f, bulk stores %.0f, release %.0f]; metadata warp [buffer wait %.0f]" % (
        *(p[:, i].sum() / tiles for i in (8, 9, 10, 11, 16, 17, 18, 20)),))
    nl = max(p[:, 21].sum(), 1)
    print("prof: consumer thread 0, cycles per own level: sweep %.0f, fence + arrivals (level barrier, stages) %.0f, load of the next own level %.0f; own levels %d" % (
        p[:, 12].sum() / nl, p[:, 15].sum() / nl, p[:, 13].sum() / nl, nl))
