"""Second, independent numpy readings of the two ends of the outer iteration against the oracle (same purpose as
tests/test_flux_second_reading.py): the residual sources of residualsUpdate.H:1-15,44-83 (steady and dual-time Euler) from the
face fluxes, and the field update of updateFields.H:1-104 from the solved increments."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle


def surface_integrate(mesh, flux):
    """fvc::surfaceIntegrate without the division by V: owner += flux, neighbour -= flux, boundary faces into their cell"""
    F = mesh.n_internal_faces
    out = np.zeros((mesh.n_cells,) + flux.shape[1:])
    np.add.at(out, mesh.owner[:F], flux[:F])
    np.subtract.at(out, mesh.neighbour, flux[:F])
    np.add.at(out, mesh.owner[F:], flux[F:])
    return out


@pytest.mark.parametrize("make", [lambda: cases.onera_box(6), lambda: cases.bump(12, 9), lambda: cases.periodic_box(6, "AUSMPlusUp", "vanLeer", seed=8)])
def test_steady_sources_are_minus_the_surface_integral_of_the_fluxes(make):
    case = make()
    o = case.apply(Oracle())
    phi, phiUp, phiEp = o.calc_flux()
    src = o.residual()
    for got, flux in zip(src, (phi, phiUp, phiEp)):
        want = -surface_integrate(case.mesh, flux)
        assert np.abs(got - want).max() <= 1e-12 * np.abs(flux).max()


def test_dual_time_euler_sources():
    """transient: R V -= (ddt.diag * W - ddt.source) with the Euler coefficients diag = V / deltaT, source = V W.old / deltaT"""
    case = cases.shock_tube(40, "ROE")                       # Euler, deltaT from the case
    o = case.apply(Oracle())
    ctl = case.controls
    o.new_time_step()
    w_old = o.state_get()
    o.iterate(ctl)                                           # state moves away from the old time level
    fl = o.calc_flux()
    src = o.residual()
    w = o.state_get()
    dt = case.schemes.delta_t
    for got, flux, key in zip(src, fl, ("rho", "rhoU", "rhoE")):
        V = case.mesh.V if w[key].ndim == 1 else case.mesh.V[:, None]
        want = -surface_integrate(case.mesh, flux) - V * (w[key] - w_old[key]) / dt
        assert np.abs(got - want).max() <= 1e-11 * max(np.abs(flux).max(), np.abs(want).max()), key


@pytest.mark.parametrize("make", [lambda: cases.onera_box(6), lambda: cases.periodic_box(6, "ROE", "Minmod", seed=9)])
def test_update_fields_second_reading(make):
    case = make()
    o = case.apply(Oracle())
    o.calc_flux(); o.residual(); o.pseudo_dt(); o.assemble()
    before = o.state_get()
    (dr, dru, dre), _ = o.solve_delta(case.controls)
    o.update_fields()
    after = o.state_get()
    R, Cp = case.R, case.Cp
    Cv = Cp - R
    rho = before["rho"] + dr
    rhoU = before["rhoU"] + dru
    rhoE = before["rhoE"] + dre
    U = rhoU / rho[:, None]
    e = rhoE / rho - 0.5 * (U * U).sum(1)
    T = e / Cv                                               # hePsiThermo<hConst, perfectGas>, sensibleInternalEnergy, Tref = 0
    p = rho / (1.0 / (R * T))                                # rho / psi
    assert np.abs(dr).max() > 0
    for key, want in (("rho", rho), ("U", U), ("T", T), ("p", p), ("rhoU", rho[:, None] * U), ("rhoE", rho * (e + 0.5 * (U * U).sum(1)))):
        assert np.abs(after[key] - want).max() <= 1e-13 * np.abs(want).max(), key


def test_dual_time_backward_sources_and_ddt_coefficient():
    """second physical step of `backward`: R V -= V (1.5 W - 2 W.old + 0.5 W.oldOld) / deltaT (OpenFOAM backwardDdtScheme with
    deltaT = deltaT0), and the matrix diagonal carries ddtCoeff V = (rPseudoDeltaT + 1.5 / deltaT) V (dualTimeDdtScheme.C:111-126,
    outerLoop.H:61-64)"""
    case = cases.shock_tube(40, "ROE")
    case.schemes.ddt_scheme = capi.DDT_NAMES["backward"]
    o = case.apply(Oracle())
    ctl = case.controls
    o.new_time_step()
    w00 = o.state_get()
    for _ in range(3):
        o.iterate(ctl)
    o.new_time_step()
    w0 = o.state_get()
    o.iterate(ctl)
    fl = o.calc_flux()
    src = o.residual()
    w = o.state_get()
    dt = case.schemes.delta_t
    for got, flux, key in zip(src, fl, ("rho", "rhoU", "rhoE")):
        V = case.mesh.V if w[key].ndim == 1 else case.mesh.V[:, None]
        want = -surface_integrate(case.mesh, flux) - V * (1.5 * w[key] - 2.0 * w0[key] + 0.5 * w00[key]) / dt
        assert np.abs(got - want).max() <= 1e-10 * max(np.abs(flux).max(), np.abs(want).max()), key
    rdt, _ = o.pseudo_dt()
    o.assemble()
    d_with, _, _ = o.matrix_get_ldu(0)
    # the same state assembled as a steady problem differs on the diagonal by exactly 1.5 V / deltaT
    steady = cases.shock_tube(40, "ROE")
    steady.schemes.ddt_scheme = capi.DDT_NAMES["steadyState"]
    steady.p, steady.U, steady.T = w["p"], w["U"], w["T"]
    o2 = steady.apply(Oracle())
    o2.calc_flux(); o2.residual(); rdt2, _ = o2.pseudo_dt(); o2.assemble()
    d_steady, _, _ = o2.matrix_get_ldu(0)
    extra = (d_with - d_steady)[:, 0] / case.mesh.V - (rdt - rdt2)      # the pseudo-Courant numbers differ (SER history), the rest must not
    assert np.allclose(extra, 1.5 / dt, rtol=1e-9)
