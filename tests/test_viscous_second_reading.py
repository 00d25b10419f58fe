"""A second, independent numpy reading of the laminar viscous residual (residualsUpdate.H:16-43) against the oracle, on the cells of
an orthogonal box that are two layers away from every boundary (their stencil — Gauss gradients of the neighbours included —
sees cell values only, and the non-orthogonal correction of `Gauss linear corrected` vanishes):
tauMC = mu dev2(T(grad U)),  rhoUR += laplacian(mu, U) + div(tauMC),
rhoER += div(((mu grad U)_f + tauMC_f) & U_f & Sf) + laplacian(alphaEff, eCalc)  with alphaEff = gamma mu / Pr (heThermo::alphaEff for an
internal-energy thermo).
Same purpose as tests/test_flux_second_reading.py."""
import numpy as np
import pytest

from icsfoam_b200 import cases
from oracle.pyoracle import Oracle


def viscous_sources(mesh, st, mu, alpha):
    F, N = mesh.n_internal_faces, mesh.n_cells
    own, nei, w = mesh.owner[:F], mesh.neighbour, mesh.weights[:F]
    Sf, magSf, dc, V = mesh.Sf[:F], mesh.magSf[:F], mesh.nonOrthDeltaCoeffs[:F], mesh.V
    U = st["U"]
    lin = lambda a: (w.reshape((-1,) + (1,) * (a.ndim - 1))) * a[own] + ((1 - w).reshape((-1,) + (1,) * (a.ndim - 1))) * a[nei]

    def integrate(ff):                                   # sum of outward face values (internal faces only: deep-interior cells)
        out = np.zeros((N,) + ff.shape[1:])
        np.add.at(out, own, ff)
        np.subtract.at(out, nei, ff)
        return out

    gradU = integrate(Sf[:, :, None] * lin(U)[:, None, :]) / V[:, None, None]          # (grad U)_ij = d_i U_j
    tr = np.trace(gradU, axis1=1, axis2=2)
    tauMC = mu * (np.swapaxes(gradU, 1, 2) - (2.0 / 3.0) * tr[:, None, None] * np.eye(3))
    lapU = integrate((mu * magSf * dc)[:, None] * (U[nei] - U[own]))
    div_tau = integrate(np.einsum("fi,fij->fj", Sf, lin(tauMC)))
    rhoUR = lapU + div_tau
    sigma = mu * lin(gradU) + lin(tauMC)
    sigma_dot_u = np.einsum("fij,fj->fi", sigma, lin(U))
    e = st["rhoE"] / st["rho"] - 0.5 * (U * U).sum(1)
    rhoER = integrate((sigma_dot_u * Sf).sum(1)) + integrate(alpha * magSf * dc * (e[nei] - e[own]))
    return rhoUR, rhoER


@pytest.mark.parametrize("flux,seed", [("ROE", 31), ("HLLC", 32)])
def test_viscous_residual_second_reading(flux, seed):
    mu, Pr = 0.07, 0.9
    visc = cases.periodic_box(8, flux, "vanLeer", seed=seed, mu=mu, Pr=Pr)
    invisc = cases.periodic_box(8, flux, "vanLeer", seed=seed)
    mesh = visc.mesh
    out = []
    for case in (visc, invisc):
        o = case.apply(Oracle())
        o.calc_flux()
        out.append((o.residual(), o.state_get()))
    (sv, st), (si, _) = out
    assert np.array_equal(sv[0], si[0])                                  # the continuity source has no viscous part
    F = mesh.n_internal_faces
    layer = np.zeros(mesh.n_cells, bool)
    layer[mesh.owner[F:]] = True
    near = layer.copy()
    near[mesh.owner[:F][layer[mesh.neighbour]]] = True
    near[mesh.neighbour[layer[mesh.owner[:F]]]] = True
    deep = np.flatnonzero(~near)
    assert len(deep) >= 27
    assert np.allclose(mesh.nonOrthDeltaCoeffs[:F], mesh.deltaCoeffs[:F], rtol=1e-12)          # orthogonal: no correction term
    gamma = visc.Cp / (visc.Cp - visc.R)
    rhoUR, rhoER = viscous_sources(mesh, st, mu, gamma * (mu / Pr))
    got_U, got_E = sv[1] - si[1], sv[2] - si[2]
    assert np.abs(got_U[deep]).max() > 0 and np.abs(got_E[deep]).max() > 0
    assert np.abs(got_U[deep] - rhoUR[deep]).max() <= 1e-10 * np.abs(sv[1]).max()
    assert np.abs(got_E[deep] - rhoER[deep]).max() <= 1e-10 * np.abs(sv[2]).max()


def viscous_sources_corrected(mesh, st, mu, alpha):
    """as viscous_sources, with the non-orthogonal correction of `Gauss linear corrected` (OpenFOAM correctedSnGrad, component-wise for
    vectors): snGrad = nonOrthDeltaCoeffs (phi_N - phi_P) + (n - delta nonOrthDeltaCoeffs) . linearInterpolate(grad phi)"""
    F, N = mesh.n_internal_faces, mesh.n_cells
    own, nei, w = mesh.owner[:F], mesh.neighbour, mesh.weights[:F]
    Sf, magSf, ndc, V = mesh.Sf[:F], mesh.magSf[:F], mesh.nonOrthDeltaCoeffs[:F], mesh.V
    n = Sf / magSf[:, None]
    corr = n - (mesh.C[nei] - mesh.C[own]) * ndc[:, None]
    U = st["U"]
    lin = lambda a: (w.reshape((-1,) + (1,) * (a.ndim - 1))) * a[own] + ((1 - w).reshape((-1,) + (1,) * (a.ndim - 1))) * a[nei]

    def integrate(ff):
        out = np.zeros((N,) + ff.shape[1:])
        np.add.at(out, own, ff)
        np.subtract.at(out, nei, ff)
        return out

    def grad(phi):                                           # Gauss linear, (N,3) for a scalar, (N,3,3) [i: derivative, j: component] for a vector
        pf = lin(phi)
        ff = Sf * pf[:, None] if phi.ndim == 1 else Sf[:, :, None] * pf[:, None, :]
        return integrate(ff) / V.reshape((-1,) + (1,) * (ff.ndim - 1))

    gradU = grad(U)
    sn_U = ndc[:, None] * (U[nei] - U[own]) + np.einsum("fi,fij->fj", corr, lin(gradU))
    tr = np.trace(gradU, axis1=1, axis2=2)
    tauMC = mu * (np.swapaxes(gradU, 1, 2) - (2.0 / 3.0) * tr[:, None, None] * np.eye(3))
    rhoUR = integrate((mu * magSf)[:, None] * sn_U) + integrate(np.einsum("fi,fij->fj", Sf, lin(tauMC)))
    sigma = mu * lin(gradU) + lin(tauMC)
    e = st["rhoE"] / st["rho"] - 0.5 * (U * U).sum(1)
    sn_e = ndc * (e[nei] - e[own]) + (corr * lin(grad(e))).sum(1)
    rhoER = integrate((np.einsum("fij,fj->fi", sigma, lin(U)) * Sf).sum(1)) + integrate(alpha * magSf * sn_e)
    return rhoUR, rhoER


def test_viscous_residual_second_reading_on_a_non_orthogonal_mesh():
    mu, Pr = 0.04, 0.75
    visc, invisc = cases.bump(15, 12, mu=mu, Pr=Pr), cases.bump(15, 12)
    mesh = visc.mesh
    rng = np.random.default_rng(41)                          # a non-trivial state: the tutorial's initial field is uniform
    for c in (visc, invisc):
        c.U = c.U * (1 + 0.05 * np.sin(7 * mesh.C[:, :1]) * np.cos(5 * mesh.C[:, 1:2])) + np.column_stack([np.zeros(mesh.n_cells), 15 * np.sin(3 * mesh.C[:, 0]), np.zeros(mesh.n_cells)])
        c.T = c.T * (1 + 0.03 * np.cos(4 * mesh.C[:, 0] + 2 * mesh.C[:, 1]))
    out = []
    for case in (visc, invisc):
        o = case.apply(Oracle())
        o.calc_flux()
        out.append((o.residual(), o.state_get()))
    (sv, st), (si, _) = out
    F = mesh.n_internal_faces
    from icsfoam_b200 import capi
    layer = np.zeros(mesh.n_cells, bool)
    for p in mesh.patches:
        if p["kind"] != capi.EMPTY:
            layer[mesh.owner[p["start"]:p["start"] + p["size"]]] = True
    near = layer.copy()
    near[mesh.owner[:F][layer[mesh.neighbour]]] = True
    near[mesh.neighbour[layer[mesh.owner[:F]]]] = True
    deep = np.flatnonzero(~near)
    assert len(deep) >= 100
    assert not np.allclose(mesh.nonOrthDeltaCoeffs[:F], mesh.deltaCoeffs[:F], rtol=1e-6)      # the correction term is active
    gamma = visc.Cp / (visc.Cp - visc.R)
    rhoUR, rhoER = viscous_sources_corrected(mesh, st, mu, gamma * (mu / Pr))
    plainU, _ = viscous_sources(mesh, st, mu, gamma * (mu / Pr))
    got_U, got_E = sv[1] - si[1], sv[2] - si[2]
    assert np.abs(got_U[deep][:, :2] - rhoUR[deep][:, :2]).max() <= 1e-9 * np.abs(got_U[deep]).max()
    assert np.abs(got_E[deep] - rhoER[deep]).max() <= 1e-9 * np.abs(got_E[deep]).max()
    assert np.abs(got_U[deep][:, :2] - plainU[deep][:, :2]).max() > 1e-6 * np.abs(got_U[deep]).max()   # ... and matters


def test_viscous_residual_second_reading_with_walls():
    """Every cell of an orthogonal box with an isothermal no-slip wall (fixedValue U and T) and zeroGradient elsewhere: boundary-face
    contributions restated from OpenFOAM's own rules — fixedValue snGrad = deltaCoeffs (phi_b - phi_P); the patch values of a Gauss
    gradient after gaussGrad::correctBoundaryConditions, grad_b = grad_P + n (snGrad_b - n . grad_P); tauMC and sigmaDotU evaluated
    from boundary values on boundary faces; eCalc = rhoE / rho - 0.5 |U|^2 as a calculated field."""
    from icsfoam_b200 import capi
    from icsfoam_b200 import meshtools as mt
    mu, Pr = 0.05, 0.72
    mesh = mt.structured(1, 6, 5, 4, 0, (0, 0, 0), (1.2, 1.0, 0.8), patch_kinds=(capi.PATCH, capi.PATCH, capi.WALL, capi.PATCH, capi.PATCH, capi.PATCH))
    rng = np.random.default_rng(51)
    N, F = mesh.n_cells, mesh.n_internal_faces
    p = 1e5 * (1 + 0.1 * rng.random(N))
    T = 300.0 * (1 + 0.1 * rng.random(N))
    U = np.column_stack([90.0 + 40 * rng.random(N), 30 * rng.standard_normal(N), 30 * rng.standard_normal(N)])
    names = [q["name"] for q in mesh.patches]
    zg = ("zeroGradient", ())
    bcs = {nm: {"p": zg, "U": zg, "T": zg} for nm in names}
    wall = names[2]
    Tw = 320.0
    bcs[wall] = {"p": zg, "U": ("fixedValue", (0.0, 0.0, 0.0)), "T": ("fixedValue", (Tw,))}
    out = []
    for m_ in (mu, 0.0):
        case = cases.Case("wall", mesh, 287.0, 1005.0, capi.default_schemes(flux_scheme="ROE"), capi.solver_controls(), bcs, p, U, T, mu=m_, Pr=Pr)
        o = case.apply(Oracle())
        o.calc_flux()
        out.append((o.residual(), o.state_get(), o.boundary_get()))
    (sv, st, bd), (si, _, _) = out
    R, Cp = 287.0, 1005.0
    Cv, gamma = Cp - R, Cp / (Cp - R)
    alpha = gamma * (mu / Pr)
    own, nei, w = mesh.owner, mesh.neighbour, mesh.weights[:F]
    Sf, magSf, dc, V = mesh.Sf, mesh.magSf, mesh.deltaCoeffs, mesh.V
    nb = Sf[F:] / magSf[F:, None]
    fc = own[F:]
    Ub, Tb = bd["U"], bd["T"]
    wallf = mesh.patch_faces(wall) - F
    assert np.abs(Ub[wallf]).max() == 0.0 and (Tb[wallf] == Tw).all()
    Uc = st["U"]
    e = st["rhoE"] / st["rho"] - 0.5 * (Uc * Uc).sum(1)
    eb = Cv * Tb                                             # he(p, T) on the patch; 0.5 |U_b|^2 cancels in eCalc
    lin = lambda a: (w.reshape((-1,) + (1,) * (a.ndim - 1))) * a[own[:F]] + ((1 - w).reshape((-1,) + (1,) * (a.ndim - 1))) * a[nei]

    def integrate(fi, fb):
        acc = np.zeros((N,) + fi.shape[1:])
        np.add.at(acc, own[:F], fi)
        np.subtract.at(acc, nei, fi)
        np.add.at(acc, fc, fb)
        return acc

    gradU = integrate(Sf[:F, :, None] * lin(Uc)[:, None, :], Sf[F:, :, None] * Ub[:, None, :]) / V[:, None, None]
    snU_b = dc[F:, None] * (Ub - Uc[fc])                     # zeroGradient faces: U_b = U_P -> 0
    gP = gradU[fc]
    gradU_b = gP + nb[:, :, None] * (snU_b - np.einsum("fi,fij->fj", nb, gP))[:, None, :]
    dev2T = lambda g: np.swapaxes(g, 1, 2) - (2.0 / 3.0) * np.trace(g, axis1=1, axis2=2)[:, None, None] * np.eye(3)
    tau, tau_b = mu * dev2T(gradU), mu * dev2T(gradU_b)
    lapU = integrate((mu * magSf[:F] * dc[:F])[:, None] * (Uc[nei] - Uc[own[:F]]), (mu * magSf[F:])[:, None] * snU_b)
    divTau = integrate(np.einsum("fi,fij->fj", Sf[:F], lin(tau)), np.einsum("fi,fij->fj", Sf[F:], tau_b))
    sig_i = np.einsum("fij,fj->fi", mu * lin(gradU) + lin(tau), lin(Uc))
    sig_b = np.einsum("fij,fj->fi", mu * gradU_b + tau_b, Ub)
    rhoER = integrate((sig_i * Sf[:F]).sum(1), (sig_b * Sf[F:]).sum(1))
    rhoER = rhoER + integrate(alpha * magSf[:F] * dc[:F] * (e[nei] - e[own[:F]]), alpha * magSf[F:] * dc[F:] * (eb - e[fc]))
    rhoUR = lapU + divTau
    got_U, got_E = sv[1] - si[1], sv[2] - si[2]
    assert np.abs(got_U - rhoUR).max() <= 1e-10 * np.abs(got_U).max()
    assert np.abs(got_E - rhoER).max() <= 1e-10 * np.abs(got_E).max()
    wall_heat = alpha * magSf[F:][wallf] * dc[F:][wallf] * (eb[wallf] - e[fc[wallf]])
    assert np.abs(wall_heat).max() > 1e-2 * np.abs(got_E).max()                 # the wall terms are a visible part of what was compared
    assert np.abs((mu * magSf[F:])[wallf, None] * snU_b[wallf]).max() > 1e-2 * np.abs(got_U).max()
