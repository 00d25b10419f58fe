"""A second, independent reading of the reference's three Riemann solvers, in numpy, against the C++ oracle.

The oracle (oracle/oracle_flux.cpp) is a transcription of hllcFluxScheme.C:70-240, roeFluxScheme.C:41-241,276-409 and
ausmPlusUpFluxScheme.C:73-299; nothing in this image can run the reference itself ("parity unpinned").  The
functions below were written from the same reference files a second time, in whole-field numpy form that mirrors the
reference's own field expressions (outer products, `&`, `^`, pos/neg switches, surfaceFieldSelect), and only share the limited
L/R face states with the oracle (`orc_debug_reconstruct`, pinned separately by the limiter known answers).  A transcription
slip in either reading shows up as a disagreement; both being wrong in the same way is what this cannot exclude.
Bar: 1e-12 of the largest flux (different association of the same operations, no cancellation-prone reformulation)."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle import pyoracle
from oracle.pyoracle import Oracle

SMALL, VSMALL = 1e-15, 1e-300


def face_states(o, case, lim):
    """Cell fields as the flux schemes build them, and their limited L/R face values on all faces."""
    st = o.state_get()
    R, Cp = case.R, case.Cp
    gamma = Cp / (Cp - R)
    rho, p, U, T = st["rho"], st["p"], st["U"], st["T"]
    psi = 1.0 / (R * T)
    E = st["rhoE"] / rho                                          # thermo.he(p, T) + 0.5 magSqr(U)
    H = np.maximum(E, SMALL) + np.maximum(p / rho, SMALL)
    c = np.maximum(np.sqrt(gamma / psi), VSMALL)
    rec = lambda f: pyoracle.debug_reconstruct(o, lim, f)
    out = {"gamma": gamma}
    for name, f in (("rho", rho), ("p", p), ("E", E), ("H", H), ("c", c), ("cCrit", np.sqrt(2.0 * (gamma - 1.0) / (gamma + 1.0) * H))):
        out[name + "_l"], out[name + "_r"] = rec(f)
    UL, UR = np.zeros((o.mesh.n_faces, 3)), np.zeros((o.mesh.n_faces, 3))
    for k in range(3):
        UL[:, k], UR[:, k] = rec(np.ascontiguousarray(U[:, k]))
    out["U_l"], out["U_r"] = UL, UR
    return out


def rusanov(s, Sf, magSf, mrf=0.0):
    """Local Lax-Friedrichs flux of the limited states in the relative frame — this library's own scheme (no reference file): written
    here from its definition, F = (F(W_l) + F(W_r))/2 . Sf - lambda/2 (W_r - W_l) |Sf|, lambda = max(|u_l| + c_l, |u_r| + c_r),
    u = U.n - MRFFaceVelocity, energy flux rho H u + p u_mrf."""
    n = Sf / magSf[:, None]
    dot = lambda a, b: (a * b).sum(1)
    V = lambda a: a[:, None]
    u_l, u_r = dot(s["U_l"], n) - mrf, dot(s["U_r"], n) - mrf
    lam = np.maximum(np.abs(u_l) + s["c_l"], np.abs(u_r) + s["c_r"])
    W_l = (s["rho_l"], V(s["rho_l"]) * s["U_l"], s["rho_l"] * s["E_l"])
    W_r = (s["rho_r"], V(s["rho_r"]) * s["U_r"], s["rho_r"] * s["E_r"])
    F_l = (s["rho_l"] * u_l, V(s["rho_l"] * u_l) * s["U_l"] + V(s["p_l"]) * n, s["rho_l"] * u_l * s["H_l"] + s["p_l"] * mrf)
    F_r = (s["rho_r"] * u_r, V(s["rho_r"] * u_r) * s["U_r"] + V(s["p_r"]) * n, s["rho_r"] * u_r * s["H_r"] + s["p_r"] * mrf)
    phi = (0.5 * (F_l[0] + F_r[0]) - 0.5 * lam * (W_r[0] - W_l[0])) * magSf
    phiUp = (0.5 * (F_l[1] + F_r[1]) - V(0.5 * lam) * (W_r[1] - W_l[1])) * V(magSf)
    phiEp = (0.5 * (F_l[2] + F_r[2]) - 0.5 * lam * (W_r[2] - W_l[2])) * magSf
    return phi, phiUp, phiEp


def hllc(s, Sf, magSf, mrf=0.0):
    """hllcFluxScheme.C:70-240 (static mesh; mrf = MRFFaceVelocity per face)"""
    n = Sf / magSf[:, None]
    dot = lambda a, b: (a * b).sum(1)
    rho_l, rho_r, p_l, p_r, U_l, U_r = s["rho_l"], s["rho_r"], s["p_l"], s["p_r"], s["U_l"], s["U_r"]
    c_l, c_r, E_l, E_r, H_l, H_r = s["c_l"], s["c_r"], s["E_l"], s["E_r"], s["H_l"], s["H_r"]
    coefR = np.sqrt(np.maximum(VSMALL, rho_r) / np.maximum(VSMALL, rho_l))
    uAvg = (coefR[:, None] * U_r + U_l) / (coefR + 1.0)[:, None]
    HAvg = (coefR * H_r + H_l) / (coefR + 1.0)
    cAvg = np.sqrt(np.abs((s["gamma"] - 1.0) * (HAvg - 0.5 * dot(uAvg, uAvg))))
    uMag_l, uMag_r, uMagAvg = dot(U_l, n) - mrf, dot(U_r, n) - mrf, dot(uAvg, n) - mrf
    Sl = np.minimum(uMag_l - c_l, uMagAvg - cAvg)
    Sr = np.maximum(uMag_r + c_r, uMagAvg + cAvg)
    Sm = (rho_r * uMag_r * (Sr - uMag_r) - rho_l * uMag_l * (Sl - uMag_l) + p_l - p_r) / (rho_r * (Sr - uMag_r) - rho_l * (Sl - uMag_l))
    pos = lambda x: (x >= 0).astype(float)
    neg = lambda x: (x < 0).astype(float)
    coefSl, coefSr, coefSm = pos(Sl), neg(Sr), pos(Sm)
    coefSlm = (1.0 - coefSl) * coefSm
    coefSmr = (1.0 - coefSm) * (1.0 - coefSr)
    fluxRhoStar_l = Sm / (Sl - Sm) * ((Sl - uMag_l) * rho_l)
    fluxRhoStar_r = Sm / (Sr - Sm) * ((Sr - uMag_r) * rho_r)
    phi = (coefSl * rho_l * uMag_l + coefSlm * fluxRhoStar_l + coefSmr * fluxRhoStar_r + coefSr * rho_r * uMag_r) * magSf
    pStar_l = rho_l * (uMag_l - Sl) * (uMag_l - Sm) + p_l
    pStar_r = rho_r * (uMag_r - Sr) * (uMag_r - Sm) + p_r
    V = lambda a: a[:, None]
    rhoUStar_l = V(1.0 / (Sl - Sm)) * (V((Sl - uMag_l) * rho_l) * U_l + V(pStar_l - p_l) * n)
    rhoUStar_r = V(1.0 / (Sr - Sm)) * (V((Sr - uMag_r) * rho_r) * U_r + V(pStar_r - p_r) * n)
    fluxRhoUStar_l = V(Sm) * rhoUStar_l + V(pStar_l) * n
    fluxRhoUStar_r = V(Sm) * rhoUStar_r + V(pStar_r) * n
    phiUp = (V(coefSl) * (V(rho_l * uMag_l) * U_l + V(p_l) * n) + V(coefSlm) * fluxRhoUStar_l + V(coefSmr) * fluxRhoUStar_r
             + V(coefSr) * (V(rho_r * uMag_r) * U_r + V(p_r) * n)) * V(magSf)
    rhoEStar_l = 1.0 / (Sl - Sm) * ((Sl - uMag_l) * (rho_l * E_l) - p_l * uMag_l + pStar_l * Sm)
    rhoEStar_r = 1.0 / (Sr - Sm) * ((Sr - uMag_r) * (rho_r * E_r) - p_r * uMag_r + pStar_r * Sm)
    fluxRhoEStar_l = Sm * (rhoEStar_l + pStar_l) + pStar_l * mrf
    fluxRhoEStar_r = Sm * (rhoEStar_r + pStar_r) + pStar_r * mrf
    phiEp = (coefSl * rho_l * H_l * uMag_l + coefSlm * fluxRhoEStar_l + coefSmr * fluxRhoEStar_r + coefSr * rho_r * H_r * uMag_r) * magSf
    return phi, phiUp, phiEp


def roe(s, Sf, magSf, entropy_fix, mrf=0.0):
    """roeFluxScheme.C:276-409 with getRoeDissipation :41-241 (P |Lambda| P^-1 assembled block by block as there)"""
    n = Sf / magSf[:, None]
    dot = lambda a, b: (a * b).sum(1)
    V = lambda a: a[:, None]
    T = lambda a: a[:, None, None]
    outer = lambda a, b: a[:, :, None] * b[:, None, :]
    rho_l, rho_r, p_l, p_r, U_l, U_r = s["rho_l"], s["rho_r"], s["p_l"], s["p_r"], s["U_l"], s["U_r"]
    E_l, E_r, H_l, H_r = s["E_l"], s["E_r"], s["H_l"], s["H_r"]
    coefR = np.sqrt(np.maximum(VSMALL, rho_r) / np.maximum(VSMALL, rho_l))
    rhoT = coefR * rho_l
    u = (V(coefR) * U_r + U_l) / V(coefR + 1.0)
    Hroe = (coefR * H_r + H_l) / (coefR + 1.0)
    a1 = s["gamma"] - 1.0
    c = np.sqrt(np.abs(a1 * (Hroe - 0.5 * dot(u, u))))
    uMag_l, uMag_r, un = dot(U_l, n), dot(U_r, n), dot(u, n) - mrf             # only uProjRoe is made relative (roeFluxScheme.C:355-367)
    # getRoeDissipation
    theta = 0.5 * a1 * dot(u, u)
    c2 = c * c
    a2 = 1.0 / (rhoT * c * np.sqrt(2.0))
    a3 = rhoT / (c * np.sqrt(2.0))
    a4 = (theta + c2) / a1
    a5 = 1.0 - theta / c2
    a6 = theta / a1
    # the tensor with entries +-n_k / rho (or * rho) added to the dyads: components 1,2,3,5,6,7 of a row-major 3x3
    def skew(v):                       # [[0, v2, -v1], [-v2, 0, v0], [v1, -v0, 0]]
        K = np.zeros((len(v), 3, 3))
        K[:, 0, 1], K[:, 0, 2] = v[:, 2], -v[:, 1]
        K[:, 1, 0], K[:, 1, 2] = -v[:, 2], v[:, 0]
        K[:, 2, 0], K[:, 2, 1] = v[:, 1], -v[:, 0]
        return K
    invP11 = n * V(a5) - np.cross(u, n) / V(rhoT)
    invP12 = T(a1 / c2) * outer(n, u) + skew(n / V(rhoT))
    invP13 = V(-a1 / c2) * n
    invP21 = a2 * (theta - c * un)
    invP22 = -V(a2) * (a1 * u - V(c) * n)
    invP23 = a1 * a2
    invP31 = a2 * (theta + c * un)
    invP32 = -V(a2) * (a1 * u + V(c) * n)
    invP33 = a1 * a2
    L1, L2, L3 = np.abs(un), np.abs(un + c), np.abs(un - c)
    eps = entropy_fix * np.maximum(L2, L3)
    fix = lambda L: np.where(L < eps, (L * L + eps * eps) / (2.0 * eps), L)
    L1, L2, L3 = fix(L1), fix(L2), fix(L3)
    P11, P12, P13 = n, a3, a3
    P21 = outer(u, n) - skew(n * V(rhoT))
    P22 = V(a3) * (u + V(c) * n)
    P23 = V(a3) * (u - V(c) * n)
    P31 = n * V(a6) + V(rhoT) * np.cross(u, n)
    P32 = a3 * (a4 + c * un)
    P33 = a3 * (a4 - c * un)
    vT = lambda v, M: np.einsum("fi,fij->fj", v, M)            # vector & tensor
    Tv = lambda M, v: np.einsum("fij,fj->fi", M, v)            # tensor & vector
    dCR = dot(P11 * V(L1), invP11) + L2 * P12 * invP21 + L3 * P13 * invP31
    dCU = vT(P11 * V(L1), invP12) + V(L2 * P12) * invP22 + V(L3 * P13) * invP32
    dCE = dot(P11 * V(L1), invP13) + L2 * P12 * invP23 + L3 * P13 * invP33
    dMR = Tv(P21 * T(L1), invP11) + V(L2 * invP21) * P22 + V(L3 * invP31) * P23
    dMU = np.einsum("fik,fkj->fij", P21 * T(L1), invP12) + T(L2) * outer(P22, invP22) + T(L3) * outer(P23, invP32)
    dME = Tv(P21 * T(L1), invP13) + V(L2 * invP23) * P22 + V(L3 * invP33) * P23
    dER = dot(P31 * V(L1), invP11) + L2 * P32 * invP21 + L3 * P33 * invP31
    dEU = vT(P31 * V(L1), invP12) + V(L2 * P32) * invP22 + V(L3 * P33) * invP32
    dEE = dot(P31 * V(L1), invP13) + L2 * P32 * invP23 + L3 * P33 * invP33
    rhoU_l, rhoU_r = V(rho_l) * U_l, V(rho_r) * U_r
    rhoE_l, rhoE_r = rho_l * E_l, rho_r * E_r
    dRho, dRhoU, dRhoE = rho_r - rho_l, rhoU_r - rhoU_l, rhoE_r - rhoE_l
    phi = -0.5 * magSf * (dCR * dRho + dot(dCU, dRhoU) + dCE * dRhoE)
    phiUp = -0.5 * V(magSf) * (dMR * V(dRho) + Tv(dMU, dRhoU) + dME * V(dRhoE))
    phiEp = -0.5 * magSf * (dER * dRho + dot(dEU, dRhoU) + dEE * dRhoE)
    rn_l, rn_r = rho_l * uMag_l, rho_r * uMag_r
    phi = phi + 0.5 * magSf * (rn_l + rn_r)
    phiUp = phiUp + 0.5 * V(magSf) * (V(rn_l) * U_l + V(rn_r) * U_r + n * V(p_l + p_r))
    phiEp = phiEp + 0.5 * magSf * (rn_l * H_l + rn_r * H_r)
    phi = phi - 0.5 * magSf * mrf * (rho_l + rho_r)
    phiUp = phiUp - 0.5 * V(magSf * mrf) * (rhoU_l + rhoU_r)
    phiEp = phiEp - 0.5 * magSf * mrf * (rhoE_l + rhoE_r)
    return phi, phiUp, phiEp


def ausm_plus_up(s, Sf, magSf, low_mach, mrf=0.0):
    """ausmPlusUpFluxScheme.C:73-299 (beta = 1/8, alpha = 3/16, Kp = 0.25, Ku = 0.25 as coded there)"""
    dot = lambda a, b: (a * b).sum(1)
    V = lambda a: a[:, None]
    U_L, U_R = s["U_l"], s["U_r"]
    un_L, un_R = dot(U_L, Sf) / magSf - mrf, dot(U_R, Sf) / magSf - mrf
    c_L, c_R = s["cCrit_l"], s["cCrit_r"]
    c_L = c_L * c_L / np.maximum(c_L, un_L)
    c_R = c_R * c_R / np.maximum(c_R, -un_R)
    c_f = np.minimum(c_L, c_R)
    M_L, M_R = un_L / c_f, un_R / c_f
    p2 = lambda M: 0.25 * (M + 1) ** 2
    m2 = lambda M: -0.25 * (M - 1) ** 2
    sub_L, sub_R = np.abs(M_L) < 1.0, np.abs(M_R) < 1.0
    Mp_L = np.where(sub_L, p2(M_L) * (1 - 2 * m2(M_L)), np.maximum(M_L, 0))
    pp_L = np.where(sub_L, p2(M_L) * (2 - M_L - 3 * M_L * m2(M_L)), np.where(M_L > 0, 1.0, 0.0))
    Mm_R = np.where(sub_R, m2(M_R) * (1 + 2 * p2(M_R)), np.minimum(M_R, 0))
    pm_R = np.where(sub_R, m2(M_R) * (-2 - M_R + 3 * M_R * p2(M_R)), np.where(M_R < 0, 1.0, 0.0))
    p_L, p_R, rho_L, rho_R = s["p_l"], s["p_r"], s["rho_l"], s["rho_r"]
    M12 = Mp_L + Mm_R
    p12 = pp_L * p_L + pm_R * p_R
    M_mean = 0.5 * (un_L ** 2 + un_R ** 2) / c_f ** 2
    MDiff = -0.25 * np.maximum(1.0 - M_mean, 0.0) * (p_R - p_L) / (0.5 * (rho_L + rho_R) * c_f ** 2)
    flips = ((M12 > 0.0) & (M12 + MDiff <= 0.0)) | ((M12 < 0.0) & (M12 + MDiff >= 0.0))
    M12 = M12 + np.where(flips, 0.2 * MDiff, MDiff)
    if low_mach:
        p12 = p12 + (-0.25 * pp_L * pm_R * (rho_L + rho_R) * c_f * (un_R - un_L))
    left = M12 >= 0
    rhoa = M12 * c_f * np.where(left, rho_L, rho_R)
    phi = rhoa * magSf
    phiUp = V(rhoa) * np.where(V(left), U_L, U_R) * V(magSf) + V(p12) * Sf
    phiEp = rhoa * np.where(left, s["H_l"], s["H_r"]) * magSf + p12 * mrf * magSf
    return phi, phiUp, phiEp


def interior_faces(mesh):
    """internal faces whose two cells touch no boundary face: their limited states depend on cell values only"""
    F = mesh.n_internal_faces
    touches = np.zeros(mesh.n_cells, bool)
    touches[mesh.owner[F:]] = True
    return np.flatnonzero(~touches[mesh.owner[:F]] & ~touches[mesh.neighbour])


@pytest.mark.parametrize("flux,limiter,seed", [("HLLC", "vanLeer", 3), ("HLLC", "Minmod", 4), ("ROE", "vanLeer", 5), ("ROE", "Minmod", 6),
                                               ("AUSMPlusUp", "vanLeer", 7), ("AUSMPlusUp", "Minmod", 8),
                                               ("Rusanov", "vanLeer", 9), ("Rusanov", "Minmod", 10)])
def test_second_reading_agrees_with_the_oracle(flux, limiter, seed):
    case = cases.periodic_box(7, flux, limiter, seed=seed)
    o = case.apply(Oracle())
    phi, phiUp, phiEp = o.calc_flux()
    s = face_states(o, case, capi.LIM_NAMES[limiter])
    mesh = case.mesh
    if flux == "HLLC":
        mine = hllc(s, mesh.Sf, mesh.magSf)
    elif flux == "ROE":
        mine = roe(s, mesh.Sf, mesh.magSf, case.schemes.entropy_fix_coeff)
    elif flux == "Rusanov":
        mine = rusanov(s, mesh.Sf, mesh.magSf)
    else:
        mine = ausm_plus_up(s, mesh.Sf, mesh.magSf, bool(case.schemes.low_mach_ausm))
    f = interior_faces(mesh)
    assert len(f) >= 100
    for a, b, name in zip(mine, (phi, phiUp, phiEp), ("phi", "phiUp", "phiEp")):
        scale = np.abs(b[f]).max()
        assert scale > 0
        assert np.abs(a[f] - b[f]).max() <= 1e-12 * scale, (flux, name, np.abs(a[f] - b[f]).max() / scale)


def test_second_reading_covers_every_branch():
    """The random boxes above are subsonic; a supersonic and a reversed stream exercise the coefSl / coefSr / |M| >= 1 branches and
    the Harten fix is checked to be active somewhere."""
    for flux in ("HLLC", "ROE", "AUSMPlusUp"):
        for vel in ((900.0, 40.0, -30.0), (-850.0, 10.0, 20.0), (5.0, -3.0, 2.0)):
            case = cases.periodic_box(7, flux, "vanLeer", seed=11)
            case.U = case.U * 0.02 + np.array(vel)
            o = case.apply(Oracle())
            ref = o.calc_flux()
            s = face_states(o, case, capi.LIM_VANLEER)
            mesh = case.mesh
            st = o.state_get()
            mach_x = st["U"][:, 0] / np.sqrt(s["gamma"] * case.R * st["T"])
            assert (mach_x.min() > 1.5) if vel[0] > 100 else (mach_x.max() < -1.5) if vel[0] < -100 else (np.abs(mach_x).max() < 0.1)
            mine = {"HLLC": lambda: hllc(s, mesh.Sf, mesh.magSf), "ROE": lambda: roe(s, mesh.Sf, mesh.magSf, case.schemes.entropy_fix_coeff),
                    "AUSMPlusUp": lambda: ausm_plus_up(s, mesh.Sf, mesh.magSf, bool(case.schemes.low_mach_ausm))}[flux]()
            f = interior_faces(mesh)
            for a, b in zip(mine, ref):
                scale = np.abs(b[f]).max()
                assert np.abs(a[f] - b[f]).max() <= 1e-12 * scale, (flux, vel)


@pytest.mark.parametrize("flux", ["HLLC", "ROE", "AUSMPlusUp", "Rusanov"])
def test_second_reading_in_a_rotating_and_translating_frame(flux):
    """MRFFaceVelocity enters hllcFluxScheme.C:157-161,217-218, roeFluxScheme.C:366-367,402-408, ausmPlusUpFluxScheme.C:104-105,294"""
    case = cases.periodic_box(7, flux, "vanLeer", seed=17).with_mrf((30.0, -50.0, 80.0), (0.3, 0.5, -0.2), (20.0, 5.0, -10.0))
    o = case.apply(Oracle())
    ref = o.calc_flux()
    s = face_states(o, case, capi.LIM_VANLEER)
    mesh = case.mesh
    mrf = case.mrf_fields(mesh)[0]
    assert np.abs(mrf).max() > 10.0
    mine = {"HLLC": lambda: hllc(s, mesh.Sf, mesh.magSf, mrf), "ROE": lambda: roe(s, mesh.Sf, mesh.magSf, case.schemes.entropy_fix_coeff, mrf),
            "AUSMPlusUp": lambda: ausm_plus_up(s, mesh.Sf, mesh.magSf, bool(case.schemes.low_mach_ausm), mrf),
            "Rusanov": lambda: rusanov(s, mesh.Sf, mesh.magSf, mrf)}[flux]()
    f = interior_faces(mesh)
    for a, b in zip(mine, ref):
        scale = np.abs(b[f]).max()
        assert np.abs(a[f] - b[f]).max() <= 1e-12 * scale, flux


@pytest.mark.parametrize("flux", ["HLLC", "ROE", "AUSMPlusUp"])
def test_second_reading_on_cyclic_faces(flux):
    """faces of the translational cyclic pair: the same flux formulas with the neighbour cell across the pair as the right state"""
    case = cases.periodic_box(7, flux, "vanLeer", seed=19)
    o = case.apply(Oracle())
    ref = o.calc_flux()
    s = face_states(o, case, capi.LIM_VANLEER)
    mesh = case.mesh
    mine = {"HLLC": lambda: hllc(s, mesh.Sf, mesh.magSf), "ROE": lambda: roe(s, mesh.Sf, mesh.magSf, case.schemes.entropy_fix_coeff),
            "AUSMPlusUp": lambda: ausm_plus_up(s, mesh.Sf, mesh.magSf, bool(case.schemes.low_mach_ausm))}[flux]()
    F = mesh.n_internal_faces
    physical = np.zeros(mesh.n_cells, bool)                  # cells that touch a non-coupled, non-empty patch: their gradients see BC values
    for p in mesh.patches:
        if p["kind"] not in (capi.CYCLIC, capi.EMPTY):
            physical[mesh.owner[p["start"]:p["start"] + p["size"]]] = True
    faces = []
    for p in mesh.patches:
        if p["kind"] == capi.CYCLIC:
            q = mesh.patches[p["nbr_patch"]]
            fa = np.arange(p["start"], p["start"] + p["size"])
            fb = np.arange(q["start"], q["start"] + q["size"])
            ok = ~physical[mesh.owner[fa]] & ~physical[mesh.owner[fb]]
            faces.append(fa[ok])
    faces = np.concatenate(faces)
    assert len(faces) >= 2 * 9
    for a, b in zip(mine, ref):
        scale = np.abs(b[faces]).max()
        assert scale > 0 and np.abs(a[faces] - b[faces]).max() <= 1e-12 * scale, flux


def test_second_reading_follows_the_scheme_switches():
    """lowMachAusm false (ausmPlusUpFluxScheme.C:259-268 skipped) and another entropyFixCoeff (roeFluxScheme.C:255)"""
    for flux, attr, value in (("AUSMPlusUp", "low_mach_ausm", 0), ("ROE", "entropy_fix_coeff", 0.2)):
        case = cases.periodic_box(7, flux, "vanLeer", seed=29)
        case.U = case.U * 0.05 + np.array((8.0, -3.0, 2.0))           # low speed: both switches matter
        setattr(case.schemes, attr, value)
        o = case.apply(Oracle())
        ref = o.calc_flux()
        s = face_states(o, case, capi.LIM_VANLEER)
        mesh = case.mesh
        f = interior_faces(mesh)
        if flux == "ROE":
            mine, other = roe(s, mesh.Sf, mesh.magSf, 0.2), roe(s, mesh.Sf, mesh.magSf, 0.05)
        else:
            mine, other = ausm_plus_up(s, mesh.Sf, mesh.magSf, False), ausm_plus_up(s, mesh.Sf, mesh.magSf, True)
        for a, b in zip(mine, ref):
            assert np.abs(a[f] - b[f]).max() <= 1e-12 * np.abs(b[f]).max(), flux
        assert max(np.abs(a[f] - b[f]).max() / np.abs(b[f]).max() for a, b in zip(other, ref)) > 1e-6, flux
