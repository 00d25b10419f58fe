"""A second, independent reading of the inviscid approximate Jacobian (convectiveFluxScheme::addFluxTerms / addDissipationJacobian /
addTemporalTerms, convectiveFluxScheme.C:366-546, with blockFvMatrix::insertBlock / insertDissipationBlock,
blockFvMatrix.C:211-326) in numpy against all nine LDU sub-blocks of the oracle's coupledMatrix: upper and lower coefficients on
the faces, and the diagonals of the cells, that do not see a boundary.  Same purpose as tests/test_flux_second_reading.py."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle
from tests.test_flux_second_reading import face_states, interior_faces


def second_reading(mesh, s, st, R, Cp, rdt, mrf=None, omega=None):
    """-> {block: (diag contribution per cell, upper, lower)} with blocks numbered as icsb200_matrix_get_ldu"""
    F, N = mesh.n_internal_faces, mesh.n_cells
    own, nei = mesh.owner[:F], mesh.neighbour
    magSf = mesh.magSf[:F]
    n = mesh.Sf[:F] / magSf[:, None]
    g = s["gamma"]
    V = lambda a: a[:, None]
    T = lambda a: a[:, None, None]
    outer = lambda a, b: a[:, :, None] * b[:, None, :]
    eye = np.eye(3)[None]
    side = {}
    for tag in ("l", "r"):
        U, E = s["U_" + tag][:F], s["E_" + tag][:F]
        theta = 0.5 * (g - 1) * (U * U).sum(1)
        a1, a2 = g * E - theta, g - 1
        proj = (U * n).sum(1)
        side[tag] = {4: n, 6: n * V(theta) - U * V(proj), 8: outer(U, n) - a2 * outer(n, U) + T(proj) * eye, 7: n * a2,
                     2: proj * (theta - a1), 5: n * V(a1) - a2 * U * V(proj), 3: g * proj}
    c = np.sqrt(g * R * st["T"])
    w = mesh.weights[:F]
    mrf = np.zeros(F) if mrf is None else mrf[:F]
    lam = (w * c[own] + (1 - w) * c[nei]) + np.abs(((V(w) * st["U"][own] + V(1 - w) * st["U"][nei]) * n).sum(1) - mrf)
    out = {}
    shape = {0: (), 1: (), 2: (), 3: (), 4: (3,), 5: (3,), 6: (3,), 7: (3,), 8: (3, 3)}
    for b, sh in shape.items():
        ex = (slice(None),) + (None,) * len(sh)
        upper = np.zeros((F,) + sh)
        lower = np.zeros((F,) + sh)
        if b in side["l"]:
            upper += 0.5 * magSf[ex] * side["r"][b]
            lower += -0.5 * magSf[ex] * side["l"][b]
        diag = np.zeros((N,) + sh)
        np.subtract.at(diag, own, lower)                     # negSumDiag of the convective part
        np.subtract.at(diag, nei, upper)
        if b in (0, 3, 8):                                   # insertDissipationBlock (operator-=) and the temporal term
            d = 0.5 * magSf * lam
            dd = d[ex] * (eye if b == 8 else 1.0)
            upper -= dd
            lower -= dd
            np.add.at(diag, own, dd)
            np.add.at(diag, nei, dd)
            diag += (rdt * mesh.V)[ex] * (eye if b == 8 else 1.0)
            # fvj::div(w, MRFFaceVelocity magSf) subtracted (convectiveFluxScheme.C:477-481, blockFvOperatorsTemplates.C:146-200)
            sf = (mrf * magSf)[ex] * (eye if b == 8 else 1.0)
            upper -= sf * (1 - w)[ex]
            lower += sf * w[ex]
            np.subtract.at(diag, own, sf * w[ex])
            np.add.at(diag, nei, sf * (1 - w)[ex])
        if b == 8 and omega is not None:                     # addMRFSource: diag += V [omega x] (convectiveFluxScheme.C:123-139)
            K = np.zeros((N, 3, 3))
            K[:, 0, 1], K[:, 0, 2] = -omega[:, 2], omega[:, 1]
            K[:, 1, 0], K[:, 1, 2] = omega[:, 2], -omega[:, 0]
            K[:, 2, 0], K[:, 2, 1] = -omega[:, 1], omega[:, 0]
            diag += mesh.V[:, None, None] * K
        out[b] = (diag.reshape(N, -1), upper.reshape(F, -1), lower.reshape(F, -1))
    return out


@pytest.mark.parametrize("make", [lambda: cases.onera_box(7), lambda: cases.periodic_box(7, "ROE", "Minmod", seed=21)])
def test_jacobian_second_reading(make):
    case = make()
    o = case.apply(Oracle())
    o.calc_flux(); o.residual(); rdt, _ = o.pseudo_dt(); o.assemble()
    mesh = case.mesh
    lim = case.schemes.limiter_U
    assert case.schemes.limiter_T == lim
    s = face_states(o, case, lim)
    mine = second_reading(mesh, s, o.state_get(), case.R, case.Cp, rdt)
    f = interior_faces(mesh)
    F = mesh.n_internal_faces
    touches = np.zeros(mesh.n_cells, bool)
    touches[mesh.owner[F:]] = True
    cells = np.flatnonzero(~touches)
    assert len(f) >= 100 and len(cells) >= 100
    for b in range(9):
        d, u, l = o.matrix_get_ldu(b)
        md, mu, ml = mine[b]
        scale = max(np.abs(d).max(), np.abs(u).max(), np.abs(l).max())
        if b == 1:     # dSByS(0,1): no flux term (convectiveFluxScheme.C:424), only addBoundaryTerms touches its diagonal
            assert np.abs(u).max() == 0.0 and np.abs(l).max() == 0.0 and np.abs(d[cells]).max() == 0.0 and np.abs(d).max() > 0.0
            continue
        assert np.abs(u[f] - mu[f]).max() <= 1e-12 * scale, ("upper", b)
        assert np.abs(l[f] - ml[f]).max() <= 1e-12 * scale, ("lower", b)
        assert np.abs(d[cells] - md[cells]).max() <= 1e-12 * scale, ("diag", b)


def boundary_terms(mesh, st, R, Cp):
    """convectiveFluxScheme::boundaryJacobian + addBoundaryTerms (convectiveFluxScheme.C:47-120,219-352) for patches whose p, U, T are
    all zeroGradient (valueInternalCoeffs = 1, boundary values = cell values), plus the dissipation block's boundary part
    (blockFvMatrix.C:300-320, physical patches included).  -> {block: diag contribution (N, nc)}"""
    F, N = mesh.n_internal_faces, mesh.n_cells
    fc = mesh.owner[F:]
    Sf, magSf = mesh.Sf[F:], mesh.magSf[F:]
    cv, g = Cp - R, Cp / (Cp - R)
    rho, U, p, T, rhoE = st["rho"][fc], st["U"][fc], st["p"][fc], st["T"][fc], st["rhoE"][fc]
    V = lambda a: a[:, None]
    outer = lambda a, b: a[:, :, None] * b[:, None, :]
    one = np.ones_like(U)
    UdS = (U * Sf).sum(1)
    dCp, dCU, dCT = rho / p * UdS, V(rho) * Sf * one, -rho / T * UdS
    dMp = V(rho / p * UdS) * U + Sf
    dMU = V(rho)[:, :, None] * outer(U, Sf * one) + (rho * UdS)[:, None, None] * np.eye(3)[None] * one[:, :, None]
    dMT = V(-rho / T * UdS) * U
    dEp = rhoE / p * UdS + UdS
    dEU = Sf * one * V(rhoE + p) + V(rho * UdS) * U * one
    dET = UdS * (rho * cv - rhoE / T)
    magU2 = (U * U).sum(1)
    dPdRho, dUdRho, dTdRho = 0.5 * (g - 1) * magU2, -U / V(rho), -1.0 / (cv * rho) * (rhoE / rho - magU2)
    dPdRhoU, dUdRhoU, dTdRhoU = -(g - 1) * U, 1.0 / rho, -U / V(cv * rho)
    dPdRhoE, dTdRhoE = (g - 1) * np.ones_like(rho), 1.0 / (cv * rho)
    Tv = lambda M, v: np.einsum("fij,fj->fi", M, v)
    dot = lambda a, b: (a * b).sum(1)
    face = {
        0: dCp * dPdRho + dot(dCU, dUdRho) + dCT * dTdRho,
        4: V(dCp) * dPdRhoU + dCU * V(dUdRhoU) + V(dCT) * dTdRhoU,
        1: dCp * dPdRhoE + dCT * dTdRhoE,
        6: dMp * V(dPdRho) + Tv(dMU, dUdRho) + dMT * V(dTdRho),
        8: outer(dMp, dPdRhoU) + dMU * dUdRhoU[:, None, None] + outer(dMT, dTdRhoU),
        7: dMp * V(dPdRhoE) + dMT * V(dTdRhoE),
        2: dEp * dPdRho + dot(dEU, dUdRho) + dET * dTdRho,
        5: V(dEp) * dPdRhoU + dEU * V(dUdRhoU) + V(dET) * dTdRhoU,
        3: dEp * dPdRhoE + dET * dTdRhoE,
    }
    lam = np.sqrt(g * R * T) + np.abs(UdS / magSf)
    for b in (0, 3):
        face[b] = face[b] + 0.5 * magSf * lam
    face[8] = face[8] + (0.5 * magSf * lam)[:, None, None] * np.eye(3)[None]
    out = {}
    for b, v in face.items():
        acc = np.zeros((N,) + v.shape[1:])
        np.add.at(acc, fc, v)
        out[b] = acc.reshape(N, -1)
    # dSByS(0,1): its two terms cancel analytically for a perfect gas ((gamma - 1) rho / p = 1 / (cv T)); what is left is rounding
    out["term_scale_1"] = np.abs(dCp * dPdRhoE).max()
    return out


def test_jacobian_second_reading_with_zero_gradient_boundaries():
    """every patch zeroGradient in p, U, T: the limited states of ALL faces follow from cell values, so every cell's diagonal —
    boundary terms included — is checked."""
    from icsfoam_b200 import meshtools as mt
    mesh = mt.structured(1, 6, 5, 4, 0, (0, 0, 0), (1.2, 1.0, 0.8), patch_kinds=(capi.PATCH,) * 6)
    rng = np.random.default_rng(12)
    N = mesh.n_cells
    p = 1e5 * (1 + 0.1 * rng.random(N))
    T = 300.0 * (1 + 0.1 * rng.random(N))
    U = np.column_stack([150.0 + 40 * rng.random(N), 30 * rng.standard_normal(N), 30 * rng.standard_normal(N)])
    names = [q["name"] for q in mesh.patches]
    bcs = {nm: {"p": ("zeroGradient", ()), "U": ("zeroGradient", ()), "T": ("zeroGradient", ())} for nm in names}
    case = cases.Case("zg", mesh, 287.0, 1005.0, capi.default_schemes(flux_scheme="HLLC"), capi.solver_controls(), bcs, p, U, T)
    o = case.apply(Oracle())
    o.calc_flux(); o.residual(); rdt, _ = o.pseudo_dt(); o.assemble()
    st = o.state_get()
    s = face_states(o, case, case.schemes.limiter_U)
    mine = second_reading(mesh, s, st, case.R, case.Cp, rdt)
    bnd = boundary_terms(mesh, st, case.R, case.Cp)
    for b in range(9):
        d, u, l = o.matrix_get_ldu(b)
        md, mu, ml = mine[b]
        scale = max(np.abs(d).max(), np.abs(u).max(), np.abs(l).max())
        if b == 1:
            scale = bnd["term_scale_1"]
            assert np.abs(d).max() <= 1e-12 * scale
        assert np.abs(u - mu).max() <= 1e-12 * scale and np.abs(l - ml).max() <= 1e-12 * scale, b
        assert np.abs(d - (md + bnd[b])).max() <= 1e-12 * scale, ("diag", b, np.abs(d - (md + bnd[b])).max() / scale)


def test_jacobian_second_reading_in_a_rotating_frame():
    case = cases.periodic_box(7, "HLLC", "vanLeer", seed=23).with_mrf((30.0, -50.0, 80.0), (0.3, 0.5, -0.2), (20.0, 5.0, -10.0))
    o = case.apply(Oracle())
    o.calc_flux(); o.residual(); rdt, _ = o.pseudo_dt(); o.assemble()
    mesh = case.mesh
    s = face_states(o, case, case.schemes.limiter_U)
    fv, om = case.mrf_fields(mesh)
    mine = second_reading(mesh, s, o.state_get(), case.R, case.Cp, rdt, mrf=fv, omega=om)
    f = interior_faces(mesh)
    F = mesh.n_internal_faces
    touches = np.zeros(mesh.n_cells, bool)
    touches[mesh.owner[F:]] = True
    cells = np.flatnonzero(~touches)
    for b in (0, 2, 3, 4, 5, 6, 7, 8):
        d, u, l = o.matrix_get_ldu(b)
        md, mu, ml = mine[b]
        scale = max(np.abs(d).max(), np.abs(u).max(), np.abs(l).max())
        assert np.abs(u[f] - mu[f]).max() <= 1e-12 * scale, ("upper", b)
        assert np.abs(l[f] - ml[f]).max() <= 1e-12 * scale, ("lower", b)
        assert np.abs(d[cells] - md[cells]).max() <= 1e-12 * scale, ("diag", b)


def viscous_jacobian(mesh, st, mu, alpha, full):
    """viscousFluxScheme::addFluxTerms (viscousFluxScheme.C:218-262) with fvj::laplacian (blockFvOperatorsTemplates.C:387-450): the
    contributions SUBTRACTED from the blocks, constant muEff / alphaEff.  -> {block: (diag, upper, lower)} to add to the inviscid part"""
    F, N = mesh.n_internal_faces, mesh.n_cells
    own, nei, w = mesh.owner[:F], mesh.neighbour, mesh.weights[:F]
    rho, U = st["rho"], st["U"]
    E = st["rhoE"] / rho
    out = {}

    def lap(sf_face, vf, b, tensor_I=False):
        sf2 = sf_face * mesh.magSf[:F] * mesh.deltaCoeffs[:F]
        ex = (slice(None),) + (None,) * (vf.ndim - 1)
        upp, low = vf[nei] * sf2[ex], vf[own] * sf2[ex]
        diag = np.zeros((N,) + vf.shape[1:])
        np.subtract.at(diag, own, low)
        np.subtract.at(diag, nei, upp)
        if tensor_I:
            I = np.eye(3)[None]
            upp, low, diag = upp[:, None, None] * I, low[:, None, None] * I, diag[:, None, None] * I
        d0, u0, l0 = out.get(b, (0.0, 0.0, 0.0))
        out[b] = (d0 - diag.reshape(N, -1), u0 - upp.reshape(F, -1), l0 - low.reshape(F, -1))     # block -= laplacian

    if not full:
        rho_f = w * rho[own] + (1 - w) * rho[nei]
        lam = 0.5 * (mu + alpha) / rho_f
        one = np.ones(N)
        lap(lam, one, 0)
        lap(lam, one, 8, tensor_I=True)
        lap(lam, one, 3)
    else:
        muf, alf = np.full(F, mu), np.full(F, alpha)
        lap(muf, -U / rho[:, None], 6)
        lap(muf, 1.0 / rho, 8, tensor_I=True)
        lap(alf, -E / rho + (U * U).sum(1) / rho, 2)
        lap(alf, -U / rho[:, None], 5)
        lap(alf, 1.0 / rho, 3)
    return out


@pytest.mark.parametrize("full", [False, True])
def test_viscous_jacobian_second_reading(full):
    mu, Pr = 0.06, 0.8
    case = cases.periodic_box(7, "ROE", "vanLeer", seed=27, mu=mu, Pr=Pr)
    case.schemes.viscous_full_jacobian = int(full)
    o = case.apply(Oracle())
    o.calc_flux(); o.residual(); rdt, _ = o.pseudo_dt(); o.assemble()
    mesh = case.mesh
    st = o.state_get()
    s = face_states(o, case, case.schemes.limiter_U)
    inv = second_reading(mesh, s, st, case.R, case.Cp, rdt)
    gamma = case.Cp / (case.Cp - case.R)
    visc = viscous_jacobian(mesh, st, mu, gamma * (mu / Pr), full)
    f = interior_faces(mesh)
    F = mesh.n_internal_faces
    touches = np.zeros(mesh.n_cells, bool)
    touches[mesh.owner[F:]] = True
    cells = np.flatnonzero(~touches)
    assert np.allclose(mesh.nonOrthDeltaCoeffs[:F], mesh.deltaCoeffs[:F], rtol=1e-12)
    changed = 0
    for b in (0, 2, 3, 4, 5, 6, 7, 8):
        d, u, l = o.matrix_get_ldu(b)
        md, mu_, ml = inv[b]
        if b in visc:
            vd, vu, vl = visc[b]
            changed += int(np.abs(vu).max() > 1e-6 * np.abs(u).max())
            md, mu_, ml = md + vd, mu_ + vu, ml + vl
        scale = max(np.abs(d).max(), np.abs(u).max(), np.abs(l).max())
        assert np.abs(u[f] - mu_[f]).max() <= 1e-12 * scale, ("upper", b)
        assert np.abs(l[f] - ml[f]).max() <= 1e-12 * scale, ("lower", b)
        assert np.abs(d[cells] - md[cells]).max() <= 1e-12 * scale, ("diag", b)
    assert changed >= 3
