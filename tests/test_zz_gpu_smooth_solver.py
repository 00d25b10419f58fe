"""smoothSolverCoupled (smoother Jacobi) on the device against the oracle (smoothSolverCoupled.C:385-515,
JacobiSmoother.C:120-203).  The device evaluates a sweep as x + D^-1 (b - A x) with the SpMV and block-Jacobi kernels, the oracle
as D^-1 (b - (A - D) x): equal up to rounding, bar 1e-8 like every solver output.  Runs last on purpose (added after this round's
GPU minutes were spent; its first GPU run is the round-end suite — DESIGN.md §6)."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("make", [lambda: cases.onera_box(9), lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=5, mu=0.05),
                                  lambda: cases.bump(15, 10)])
def test_smooth_solver_matches_oracle(make, gpu_context):
    case = make()
    ctl = capi.solver_controls(solver="smoothSolverCoupled", n_sweeps=3, max_iter=12, tolerance=1e-12, rel_tol=1e-3)
    g, o = case.apply(gpu_context()), case.apply(Oracle())
    for api in (g, o):
        api.calc_flux(); api.residual(); api.pseudo_dt(); api.assemble()
    (gr, gru, gre), gres = g.solve_delta(ctl)
    (orr, oru, ore), ores = o.solve_delta(ctl)
    assert gres.n_iterations == ores.n_iterations and gres.n_iterations % 3 == 0
    for a, b in ((gr, orr), (gru, oru), (gre, ore)):
        assert np.abs(a - b).max() <= 1e-8 * np.abs(b).max()
    assert np.allclose(list(gres.s_final) + list(gres.v_final), list(ores.s_final) + list(ores.v_final), rtol=1e-7, atol=1e-14)
    # and as the solver of the outer iteration
    for _ in range(3):
        rg, ro = g.iterate(ctl), o.iterate(ctl)
        assert rg.n_iterations == ro.n_iterations
    sg, so = g.state_get(), o.state_get()
    for k in ("rho", "rhoU", "rhoE"):
        assert np.abs(sg[k] - so[k]).max() <= 1e-8 * np.abs(so[k]).max(), k


def test_smooth_solver_is_refused_for_harmonic_balance(gpu_context):
    hbcase = cases.hb_box(4, 3, flux="ROE")
    g = hbcase.apply(gpu_context())
    ctl = capi.solver_controls(solver="smoothSolverCoupled", n_sweeps=2, max_iter=4)
    with pytest.raises(capi.ApiError, match="Harmonic Balance"):
        g.iterate(ctl)
