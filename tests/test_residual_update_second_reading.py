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
