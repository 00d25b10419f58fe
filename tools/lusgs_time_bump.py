"""Development aid: LU-SGS / SpMV kernel times on the 2-D bump mesh (C3) for the level and tile schedules."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import cases
from icsfoam_b200.context import Context
nxb = int(sys.argv[1]) if len(sys.argv) > 1 else 320
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 260
case = cases.bump(nxb, ny)
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
