"""Development script: first GPU-vs-oracle comparison (superseded by tests/test_gpu_*.py)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from icsfoam_b200 import cases, capi
from icsfoam_b200.context import Context
from oracle.pyoracle import Oracle


def cmp(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(np.abs(b).max(), 1e-300)
    d = np.abs(a - b).max()
    print(f"  {name:28s} max|diff|={d:.3e} rel={d/scale:.3e} exact={np.array_equal(a, b)}")


for mk in (lambda: cases.periodic_box(6, "HLLC", "vanLeer"), lambda: cases.periodic_box(6, "ROE", "Minmod"),
           lambda: cases.periodic_box(6, "AUSMPlusUp", "vanLeer"), lambda: cases.bump(12, 8), lambda: cases.onera_box(8)):
    c = mk()
    print("case", c.name, c.mesh.n_cells, "flux", c.schemes.flux_scheme)
    o = c.apply(Oracle())
    g = c.apply(Context())
    print("  schedule", g.schedule_info())
    so, sg = o.state_get(), g.state_get()
    for k in so: cmp("state0." + k, sg[k], so[k])
    bo, bg = o.boundary_get(), g.boundary_get()
    for k in bo: cmp("bnd0." + k, bg[k], bo[k])
    fo, fg = o.calc_flux(), g.calc_flux()
    for n, a, b in zip(("phi", "phiUp", "phiEp"), fg, fo): cmp(n, a, b)
    ro, rg = o.residual(), g.residual()
    for n, a, b in zip(("srcRho", "srcRhoU", "srcRhoE"), rg, ro): cmp(n, a, b)
    do, dg = o.pseudo_dt(), g.pseudo_dt()
    cmp("rPseudoDeltaT", dg[0], do[0]); cmp("pseudoCo", dg[1], do[1])
    o.assemble(); g.assemble()
    for blk in range(9):
        lo, lg = o.matrix_get_ldu(blk), g.matrix_get_ldu(blk)
        for n, a, b in zip(("diag", "upper", "lower"), lg, lo): cmp(f"blk{blk}.{n}", a, b)
    rng = np.random.default_rng(1)
    N = c.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    if c.mesh.solutionD[2] == -1: x[1][:, 2] = 0
    yo, yg = o.matrix_mul(*x), g.matrix_mul(*x)
    for n, a, b in zip(("Ax.rho", "Ax.rhoU", "Ax.rhoE"), yg, yo): cmp(n, a, b)
    for pk in ("LUSGS", "Jacobi"):
        po, pg = o.precondition(pk, *x), g.precondition(pk, *x)
        for n, a, b in zip(("rho", "rhoU", "rhoE"), pg, po): cmp(f"{pk}.{n}", a, b)
    (dwo, reso), (dwg, resg) = o.solve_delta(c.controls), g.solve_delta(c.controls)
    for n, a, b in zip(("dRho", "dRhoU", "dRhoE"), dwg, dwo): cmp(n, a, b)
    print("  res oracle", reso.as_dict()); print("  res gpu   ", resg.as_dict())
    o.update_fields(); g.update_fields()
    so, sg = o.state_get(), g.state_get()
    for k in so: cmp("state1." + k, sg[k], so[k])
    for it in range(5):
        ro_, rg_ = o.iterate(c.controls), g.iterate(c.controls)
    print("  it5 res oracle", list(ro_.s_init), ro_.n_iterations); print("  it5 res gpu   ", list(rg_.s_init), rg_.n_iterations)
    so, sg = o.state_get(), g.state_get()
    for k in so: cmp("state6." + k, sg[k], so[k])
    g.close(); o.close()
