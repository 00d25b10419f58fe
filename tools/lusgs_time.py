"""Development aid: schedule statistics and LU-SGS / SpMV kernel times on the GPU box."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import cases
from icsfoam_b200.context import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
case = cases.onera_box(n)
g = case.apply(Context())
print("schedule", g.schedule_info())
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
