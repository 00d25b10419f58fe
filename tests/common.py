"""Shared helpers of the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def run_sequence(api, case, n_iter=3):
    """The piecewise hot path followed by n_iter fused iterations; mirrors tests/golden/make_golden.py::run."""
    out = {}
    if case.schemes.ddt_scheme != 0:
        api.new_time_step()
    phi, phiUp, phiEp = api.calc_flux()
    out.update(phi=phi, phiUp=phiUp, phiEp=phiEp)
    r = api.residual()
    out.update(srcRho=r[0], srcRhoU=r[1], srcRhoE=r[2])
    rdt, co = api.pseudo_dt()
    out.update(rPseudoDeltaT=rdt)
    api.assemble()
    for b in range(9):
        d, u, l = api.matrix_get_ldu(b)
        out[f"diag{b}"], out[f"upper{b}"], out[f"lower{b}"] = d, u, l
    (dr, dru, dre), res = api.solve_delta(case.controls)
    out.update(dRho=dr, dRhoU=dru, dRhoE=dre, restarts=np.array([res.n_iterations]), sInit=np.array(res.s_init[:]),
               vInit=np.array(res.v_init[:]))
    api.update_fields()
    hist = []
    for _ in range(n_iter):
        rr = api.iterate(case.controls)
        hist.append(list(rr.s_init) + list(rr.v_init) + [rr.n_iterations])
    st = api.state_get()
    out.update(history=np.array(hist), rho=st["rho"], rhoU=st["rhoU"], rhoE=st["rhoE"])
    return out


# tolerance per quantity: values computed without reductions are bit-comparable (1e-12 leaves room for libm pow);
# everything downstream of GMRES dot products carries reduction-order noise amplified by the Krylov recurrence
EXACT_KEYS = ["phi", "phiUp", "phiEp", "srcRho", "srcRhoU", "srcRhoE", "rPseudoDeltaT"] + [f"{k}{b}" for b in range(9) for k in ("diag", "upper", "lower")]
SOLVE_KEYS = ["dRho", "dRhoU", "dRhoE"]
STATE_KEYS = ["rho", "rhoU", "rhoE"]
TOL_EXACT = 1e-12   # north_star: per-face fluxes and assembled Jacobian blocks to 1e-12 relative
TOL_SOLVE = 1e-8
TOL_STATE = 1e-8    # north_star: converged fields within 1e-8 relative
