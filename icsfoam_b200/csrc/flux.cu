// flux.cu — Gauss gradients, limited L/R reconstruction, convective flux (HLLC / ROE / AUSM+up) and residual.
//
// Replaces the chain of whole-field fvc:: operations of convectiveFluxScheme::calcFlux
// (hllcFluxScheme.C:70-240, roeFluxScheme.C:41-241 + 276-409, ausmPlusUpFluxScheme.C:73-299) — 14-16 limited
// fvc::interpolate calls, each with its own fvc::grad and halo exchange, and ~45-60 face-field temporaries —
// and of residualsUpdate.H:1-83, by two kernels:
//   k_grad : one thread per cell row, gathers the cell's faces in ascending face id (= the order gaussGrad::gradf's
//            face loop touches that cell), 8 scalars at once;
//   k_flux_faces : one thread per cell row; reconstructs L/R states (NVD/TVD r, vanLeer / Minmod) and evaluates the flux
//            of every face the row OWNS (upper, coupled, physical) once, in the face's own orientation, 5 doubles per
//            face in GPU-face order (the fp64 pipe, not HBM, bounds this kernel: ~42 IEEE divisions/sqrts per face);
//   k_flux_gather : one thread per row; sums its faces' fluxes in ascending face id (fvc::surfaceIntegrate order; the
//            faces where it is the neighbour with a minus sign), adds the viscous divergences, applies the dual-time
//            terms (dualTimeDdtScheme.C:111-126, residualsUpdate.H:74-79) and writes the sources R*V.
// No atomics, no colouring: every output is owned by exactly one thread.  HBM-bound reads are coalesced across
// the 32 rows of a slice; neighbour gathers hit L2/L1 (hyperplane ordering keeps the three live levels resident).
// Arithmetic: fp64, -fmad=false, same operand order as the reference expressions.
#include "common.cuh"

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ double magSqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ double sqr(double x) { return x * x; }
__device__ __forceinline__ double pos0(double s) { return s >= 0 ? 1.0 : 0.0; }
__device__ __forceinline__ double negf(double s) { return s < 0 ? 1.0 : 0.0; }
__device__ __forceinline__ double sgn(double s) { return s >= 0 ? 1.0 : -1.0; }

// ---- NVD/TVD limited reconstruction of one scalar at one face (limitedScheme::calcLimiter + weights) ----
__device__ __forceinline__ double nvdR(double gradcf, double gradf)
{
    if (fabs(gradcf) >= 1000 * fabs(gradf)) return 2 * 1000 * sgn(gradcf) * sgn(gradf) - 1;
    return 2 * (gradcf / gradf) - 1;
}
__device__ __forceinline__ double limiterOf(int lim, double gradcf, double gradf)
{
    if (lim == ICSB200_LIM_VANLEER) { double r = nvdR(gradcf, gradf); return (r + fabs(r)) / (1 + fabs(r)); }
    if (lim == ICSB200_LIM_MINMOD) { double r = nvdR(gradcf, gradf); return fmax(fmin(r, 1.0), 0.0); }
    return lim == ICSB200_LIM_LINEAR ? 1.0 : 0.0;
}
// P = face owner, N = face neighbour; gP/gN their Gauss gradients; d = C_N - C_P; w = linear weight of P
__device__ __forceinline__ void reconLR(int lim, bool coupled, double phiP, double phiN, V3 gP, V3 gN, V3 d, double w, double& L, double& R)
{
    const double gradf = phiN - phiP;
    const double limL = limiterOf(lim, d.x * gP.x + d.y * gP.y + d.z * gP.z, gradf);
    const double limR = limiterOf(lim, d.x * gN.x + d.y * gN.y + d.z * gN.z, gradf);
    const double wL = limL * w + (1.0 - limL) * 1.0;
    const double wR = limR * w + (1.0 - limR) * 0.0;
    if (coupled) { L = wL * phiP + (1.0 - wL) * phiN; R = wR * phiP + (1.0 - wR) * phiN; }
    else { L = wL * (phiP - phiN) + phiN; R = wR * (phiP - phiN) + phiN; }
}

struct FaceState {
    double rho_l, rho_r, p_l, p_r, c_l, c_r, E_l, E_r, H_l, H_r;
    V3 U_l, U_r;
};
struct Flux5 { double phi; V3 phiUp; double phiEp; };

struct SchemePrm { double gamma, entropyFix; int lowMach; };

// ---- HLLC (hllcFluxScheme.C:125-238) ----
__device__ __forceinline__ Flux5 fluxHLLC(const FaceState& s, V3 Sf, double magSf, double mrf, const SchemePrm& pr)
{
    const V3 n = Sf / magSf;
    const double coefR = sqrt(fmax(ICS_VSMALL, s.rho_r) / fmax(ICS_VSMALL, s.rho_l));
    const V3 uAvg = (coefR * s.U_r + s.U_l) / (coefR + 1.0);
    const double HAvg = (coefR * s.H_r + s.H_l) / (coefR + 1.0);
    const double cAvg = sqrt(fabs((pr.gamma - 1.0) * (HAvg - 0.5 * magSqr(uAvg))));
    // contravariant velocities relative to the frame: -= MRFFaceVelocity (hllcFluxScheme.C:157-161)
    const double uMag_l = dot(s.U_l, n) - mrf, uMag_r = dot(s.U_r, n) - mrf, uMagAvg = dot(uAvg, n) - mrf;
    const double Sl = fmin(uMag_l - s.c_l, uMagAvg - cAvg);
    const double Sr = fmax(uMag_r + s.c_r, uMagAvg + cAvg);
    const double Sm = (s.rho_r * uMag_r * (Sr - uMag_r) - s.rho_l * uMag_l * (Sl - uMag_l) + s.p_l - s.p_r) /
                      (s.rho_r * (Sr - uMag_r) - s.rho_l * (Sl - uMag_l));
    const double coefSl = pos0(Sl), coefSr = negf(Sr), coefSm = pos0(Sm);
    const double coefSlm = (1.0 - coefSl) * coefSm;
    const double coefSmr = (1.0 - coefSm) * (1.0 - coefSr);
    Flux5 F;
    {
        const double starL = Sm / (Sl - Sm) * ((Sl - uMag_l) * s.rho_l);
        const double starR = Sm / (Sr - Sm) * ((Sr - uMag_r) * s.rho_r);
        F.phi = (coefSl * s.rho_l * uMag_l + coefSlm * starL + coefSmr * starR + coefSr * s.rho_r * uMag_r) * magSf;
    }
    const double pStar_l = s.rho_l * (uMag_l - Sl) * (uMag_l - Sm) + s.p_l;
    const double pStar_r = s.rho_r * (uMag_r - Sr) * (uMag_r - Sm) + s.p_r;
    {
        const V3 rhoUStar_l = 1.0 / (Sl - Sm) * ((Sl - uMag_l) * s.rho_l * s.U_l + (pStar_l - s.p_l) * n);
        const V3 rhoUStar_r = 1.0 / (Sr - Sm) * ((Sr - uMag_r) * s.rho_r * s.U_r + (pStar_r - s.p_r) * n);
        const V3 fl = Sm * rhoUStar_l + pStar_l * n;
        const V3 fr = Sm * rhoUStar_r + pStar_r * n;
        F.phiUp = (coefSl * (s.rho_l * uMag_l * s.U_l + s.p_l * n) + coefSlm * fl + coefSmr * fr + coefSr * (s.rho_r * uMag_r * s.U_r + s.p_r * n)) * magSf;
    }
    {
        const double rhoEStar_l = 1.0 / (Sl - Sm) * ((Sl - uMag_l) * (s.rho_l * s.E_l) - s.p_l * uMag_l + pStar_l * Sm);
        const double rhoEStar_r = 1.0 / (Sr - Sm) * ((Sr - uMag_r) * (s.rho_r * s.E_r) - s.p_r * uMag_r + pStar_r * Sm);
        const double fl = Sm * (rhoEStar_l + pStar_l) + pStar_l * mrf;  // hllcFluxScheme.C:217-218
        const double fr = Sm * (rhoEStar_r + pStar_r) + pStar_r * mrf;
        F.phiEp = (coefSl * s.rho_l * s.H_l * uMag_l + coefSlm * fl + coefSmr * fr + coefSr * s.rho_r * s.H_r * uMag_r) * magSf;
    }
    return F;
}

// ---- ROE (roeFluxScheme.C:41-241, 323-408) ----
struct T9 { double xx, xy, xz, yx, yy, yz, zx, zy, zz; };
__device__ __forceinline__ T9 outer(V3 a, V3 b) { return {a.x * b.x, a.x * b.y, a.x * b.z, a.y * b.x, a.y * b.y, a.y * b.z, a.z * b.x, a.z * b.y, a.z * b.z}; }
__device__ __forceinline__ T9 operator*(T9 a, double s) { return {a.xx * s, a.xy * s, a.xz * s, a.yx * s, a.yy * s, a.yz * s, a.zx * s, a.zy * s, a.zz * s}; }
__device__ __forceinline__ T9 operator+(T9 a, T9 b) { return {a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yx + b.yx, a.yy + b.yy, a.yz + b.yz, a.zx + b.zx, a.zy + b.zy, a.zz + b.zz}; }
__device__ __forceinline__ V3 mulTV(const T9& t, V3 v) { return {t.xx * v.x + t.xy * v.y + t.xz * v.z, t.yx * v.x + t.yy * v.y + t.yz * v.z, t.zx * v.x + t.zy * v.y + t.zz * v.z}; }
__device__ __forceinline__ V3 mulVT(V3 v, const T9& t) { return {v.x * t.xx + v.y * t.yx + v.z * t.zx, v.x * t.xy + v.y * t.yy + v.z * t.zy, v.x * t.xz + v.y * t.yz + v.z * t.zz}; }
__device__ __forceinline__ T9 mulTT(const T9& a, const T9& b)
{
    return {a.xx * b.xx + a.xy * b.yx + a.xz * b.zx, a.xx * b.xy + a.xy * b.yy + a.xz * b.zy, a.xx * b.xz + a.xy * b.yz + a.xz * b.zz,
            a.yx * b.xx + a.yy * b.yx + a.yz * b.zx, a.yx * b.xy + a.yy * b.yy + a.yz * b.zy, a.yx * b.xz + a.yy * b.yz + a.yz * b.zz,
            a.zx * b.xx + a.zy * b.yx + a.zz * b.zx, a.zx * b.xy + a.zy * b.yy + a.zz * b.zy, a.zx * b.xz + a.zy * b.yz + a.zz * b.zz};
}

__device__ __forceinline__ Flux5 fluxROE(const FaceState& s, V3 Sf, double magSf, double mrf, const SchemePrm& pr)
{
    const V3 n = Sf / magSf;
    const double coefR = sqrt(fmax(ICS_VSMALL, s.rho_r) / fmax(ICS_VSMALL, s.rho_l));
    const double rhoT = coefR * s.rho_l;
    const V3 uT = (coefR * s.U_r + s.U_l) / (coefR + 1.0);
    const double HT = (coefR * s.H_r + s.H_l) / (coefR + 1.0);
    const double cT = sqrt(fabs((pr.gamma - 1.0) * (HT - 0.5 * magSqr(uT))));
    const double uMag_l = dot(s.U_l, n), uMag_r = dot(s.U_r, n);
    const double uProj = dot(uT, n) - mrf;  // uProjRoe -= MRFFaceVelocity (roeFluxScheme.C:363)
    const V3 rhoU_l = s.rho_l * s.U_l, rhoU_r = s.rho_r * s.U_r;
    const double rhoE_l = s.rho_l * s.E_l, rhoE_r = s.rho_r * s.E_r;
    // eigen-decomposition in conservative variables
    const double a1 = pr.gamma - 1;
    const double theta = 0.5 * a1 * magSqr(uT);
    const double c2 = sqr(cT);
    const double a2 = 1 / (rhoT * cT * sqrt(2.0));
    const double a3 = rhoT / (cT * sqrt(2.0));
    const double a4 = (theta + c2) / a1;
    const double a5 = 1 - theta / c2;
    const double a6 = theta / a1;
    const V3 invP11 = (n * a5) - cross(uT, n) / rhoT;
    T9 invP12 = outer(a1 / c2 * n, uT);
    invP12.xy += n.z / rhoT; invP12.xz -= n.y / rhoT; invP12.yx -= n.z / rhoT;
    invP12.yz += n.x / rhoT; invP12.zx += n.y / rhoT; invP12.zy -= n.x / rhoT;
    const V3 invP13 = -a1 / c2 * n;
    const double invP21 = a2 * (theta - cT * uProj);
    const V3 invP22 = -a2 * (a1 * uT - cT * n);
    const double invP23 = a1 * a2;
    const double invP31 = a2 * (theta + cT * uProj);
    const V3 invP32 = -a2 * (a1 * uT + cT * n);
    const double invP33 = a1 * a2;
    double L1 = fabs(uProj), L2 = fabs(uProj + cT), L3 = fabs(uProj - cT);
    const double eps = pr.entropyFix * fmax(L2, L3);
    if (L1 < eps) L1 = (sqr(L1) + sqr(eps)) / (2.0 * eps);
    if (L2 < eps) L2 = (sqr(L2) + sqr(eps)) / (2.0 * eps);
    if (L3 < eps) L3 = (sqr(L3) + sqr(eps)) / (2.0 * eps);
    const V3 P11 = n;
    const double P12 = a3, P13 = a3;
    T9 P21 = outer(uT, n);
    P21.xy -= n.z * rhoT; P21.xz += n.y * rhoT; P21.yx += n.z * rhoT;
    P21.yz -= n.x * rhoT; P21.zx -= n.y * rhoT; P21.zy += n.x * rhoT;
    const V3 P22 = a3 * (uT + cT * n);
    const V3 P23 = a3 * (uT - cT * n);
    const V3 P31 = n * a6 + rhoT * cross(uT, n);
    const double P32 = a3 * (a4 + cT * uProj);
    const double P33 = a3 * (a4 - cT * uProj);
    const double dRho = s.rho_r - s.rho_l;
    const V3 dRhoU = rhoU_r - rhoU_l;
    const double dRhoE = rhoE_r - rhoE_l;
    Flux5 F;
    {
        const double dCR = dot(P11 * L1, invP11) + (L2 * P12 * invP21) + (L3 * P13 * invP31);
        const V3 dCRU = mulVT(P11 * L1, invP12) + (L2 * P12 * invP22) + (L3 * P13 * invP32);
        const double dCRE = dot(P11 * L1, invP13) + (L2 * P12 * invP23) + (L3 * P13 * invP33);
        F.phi = -0.5 * magSf * (dCR * dRho + dot(dCRU, dRhoU) + dCRE * dRhoE);
    }
    {
        const T9 P21L = P21 * L1;
        const V3 dMR = mulTV(P21L, invP11) + (L2 * P22 * invP21) + (L3 * P23 * invP31);
        const T9 dMRU = mulTT(P21L, invP12) + outer(L2 * P22, invP22) + outer(L3 * P23, invP32);
        const V3 dMRE = mulTV(P21L, invP13) + (L2 * P22 * invP23) + (L3 * P23 * invP33);
        F.phiUp = (-0.5 * magSf) * (dMR * dRho + mulTV(dMRU, dRhoU) + dMRE * dRhoE);
    }
    {
        const double dER = dot(P31 * L1, invP11) + (L2 * P32 * invP21) + (L3 * P33 * invP31);
        const V3 dERU = mulVT(P31 * L1, invP12) + (L2 * P32 * invP22) + (L3 * P33 * invP32);
        const double dERE = dot(P31 * L1, invP13) + (L2 * P32 * invP23) + (L3 * P33 * invP33);
        F.phiEp = -0.5 * magSf * (dER * dRho + dot(dERU, dRhoU) + dERE * dRhoE);
    }
    const double rhoUNorm_l = s.rho_l * uMag_l, rhoUNorm_r = s.rho_r * uMag_r;
    F.phi += 0.5 * magSf * (rhoUNorm_l + rhoUNorm_r);
    F.phiUp = F.phiUp + (0.5 * magSf) * (rhoUNorm_l * s.U_l + rhoUNorm_r * s.U_r + n * (s.p_l + s.p_r));
    F.phiEp += 0.5 * magSf * (rhoUNorm_l * s.H_l + rhoUNorm_r * s.H_r);
    // roeFluxScheme.C:406-408
    F.phi -= 0.5 * magSf * mrf * (s.rho_l + s.rho_r);
    F.phiUp = F.phiUp - (0.5 * magSf * mrf) * (rhoU_l + rhoU_r);
    F.phiEp -= 0.5 * magSf * mrf * (rhoE_l + rhoE_r);
    return F;
}

// ---- AUSM+up (ausmPlusUpFluxScheme.C:93-292) ----
__device__ __forceinline__ Flux5 fluxAUSM(const FaceState& s, V3 Sf, double magSf, double mrf, const SchemePrm& pr)
{
    const double un_L = dot(s.U_l, Sf) / magSf - mrf, un_R = dot(s.U_r, Sf) / magSf - mrf;  // ausmPlusUpFluxScheme.C:101-102
    const double c_L = sqr(s.c_l) / fmax(s.c_l, un_L);
    const double c_R = sqr(s.c_r) / fmax(s.c_r, -un_R);
    const double c_face = fmin(c_L, c_R);
    const double ML = un_L / c_face;
    double MplusL, pplusL;
    if (fabs(ML) < 1.0) {
        const double m2p = 0.25 * sqr(ML + 1), m2m = -0.25 * sqr(ML - 1);
        MplusL = m2p * (1 - 2 * m2m);
        pplusL = m2p * (2 - ML - 3 * ML * m2m);
    } else {
        MplusL = fmax(ML, 0.0);
        pplusL = (ML > 0 ? 1.0 : 0.0);
    }
    const double MR = un_R / c_face;
    double MminusR, pminusR;
    if (fabs(MR) < 1.0) {
        const double m2m = -0.25 * sqr(MR - 1), m2p = 0.25 * sqr(MR + 1);
        MminusR = m2m * (1 + 2 * m2p);
        pminusR = m2m * (-2 - MR + 3 * MR * m2p);
    } else {
        MminusR = fmin(MR, 0.0);
        pminusR = (MR < 0 ? 1.0 : 0.0);
    }
    double M12 = MplusL + MminusR;
    double p12 = pplusL * s.p_l + pminusR * s.p_r;
    const double Mmean = 0.5 * (sqr(un_L) + sqr(un_R)) / sqr(c_face);
    const double MDiff = -0.25 * fmax((1.0 - Mmean), 0.0) * (s.p_r - s.p_l) / (0.5 * (s.rho_l + s.rho_r) * sqr(c_face));
    if ((M12 > 0.0 && M12 + MDiff <= 0.0) || (M12 < 0.0 && M12 + MDiff >= 0.0)) M12 += 0.2 * MDiff;
    else M12 += MDiff;
    if (pr.lowMach) p12 += -0.25 * pplusL * pminusR * (s.rho_l + s.rho_r) * c_face * (un_R - un_L);
    const bool left = M12 >= 0;
    const double rhoa = M12 * c_face * (left ? s.rho_l : s.rho_r);
    const V3 rhoaU = rhoa * (left ? s.U_l : s.U_r);
    const double rhoah = rhoa * (left ? s.H_l : s.H_r);
    Flux5 F;
    F.phi = rhoa * magSf;
    F.phiUp = rhoaU * magSf + p12 * Sf;
    F.phiEp = rhoah * magSf + p12 * mrf * magSf;  // ausmPlusUpFluxScheme.C:293
    return F;
}

// ---- Rusanov / local Lax-Friedrichs (no reference counterpart: the flux whose frozen-lambda linearisation is the reference's
// approximate Jacobian, convectiveFluxScheme.C:402-546): central flux of the limited states in the relative frame minus
// lambda/2 (W_R - W_L), lambda = max(|u_L| + c_L, |u_R| + c_R) with u = U.n - MRFFaceVelocity.  Same expressions, same order,
// in oracle/oracle_flux.cpp fluxRusanov.
__device__ __forceinline__ Flux5 fluxRusanov(const FaceState& s, V3 Sf, double magSf, double mrf, const SchemePrm&)
{
    const V3 n = Sf / magSf;
    const double uL = dot(s.U_l, n) - mrf, uR = dot(s.U_r, n) - mrf;
    const double lam = fmax(fabs(uL) + s.c_l, fabs(uR) + s.c_r);
    const double mL = s.rho_l * uL, mR = s.rho_r * uR;
    Flux5 F;
    F.phi = (0.5 * (mL + mR) - 0.5 * lam * (s.rho_r - s.rho_l)) * magSf;
    F.phiUp = (0.5 * ((mL * s.U_l + s.p_l * n) + (mR * s.U_r + s.p_r * n)) - (0.5 * lam) * (s.rho_r * s.U_r - s.rho_l * s.U_l)) * magSf;
    F.phiEp = (0.5 * ((mL * s.H_l + s.p_l * mrf) + (mR * s.H_r + s.p_r * mrf)) - 0.5 * lam * (s.rho_r * s.E_r - s.rho_l * s.E_l)) * magSf;
    return F;
}

template <int SCHEME>
__device__ __forceinline__ Flux5 faceFlux(const FaceState& s, V3 Sf, double magSf, double mrf, const SchemePrm& pr)
{
    if (SCHEME == ICSB200_FLUX_HLLC) return fluxHLLC(s, Sf, magSf, mrf, pr);
    if (SCHEME == ICSB200_FLUX_ROE) return fluxROE(s, Sf, magSf, mrf, pr);
    if (SCHEME == ICSB200_FLUX_RUSANOV) return fluxRusanov(s, Sf, magSf, mrf, pr);
    return fluxAUSM(s, Sf, magSf, mrf, pr);
}

// ------------------------------------------------------------------------------------------------ k_grad
// gaussGrad::gradf for the NQ reconstructed scalars.  f: fields [Q_COUNT][NX]; grad: [NQ*3][NPH]
template <int NQ>
__global__ void __launch_bounds__(128, NQ > 4 ? 4 : (NQ > 1 ? 5 : 8))
k_grad(int NP, const int* __restrict__ pos2cell, const int* __restrict__ sliceOff, const int* __restrict__ rowNAll, const int* __restrict__ col,
       const int* __restrict__ meta, const int* __restrict__ gfid, const double* __restrict__ geo, size_t NFG, const double* __restrict__ V,
       const double* __restrict__ f, size_t NX, double* __restrict__ grad, size_t NPH)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)sliceOff[p >> 5];
    double own[NQ], acc[NQ][3];
#pragma unroll
    for (int k = 0; k < NQ; k++) { own[k] = f[k * NX + p]; acc[k][0] = acc[k][1] = acc[k][2] = 0.0; }
    const int nAll = rowNAll[p];
    // Software-pipelined over the row's faces: the indices and geometry of face j+1 are requested before face j is
    // accumulated, so the neighbour gathers of the next face (the kernel is bound by the latency of these L2 gathers, not by
    // bandwidth) start without waiting for an index load; the accumulation order (ascending face id) is untouched.
    struct Face { int c, type; double Sx, Sy, Sz, w; };
    auto fetch = [&](int j, Face& F) {
        const size_t e = (base + j) * 32 + lane;
        F.c = col[e];
        F.type = meta[e] & 3;
        const size_t g = gfid[e];
        F.Sx = geo[G_SFX * NFG + g]; F.Sy = geo[G_SFY * NFG + g]; F.Sz = geo[G_SFZ * NFG + g]; F.w = geo[G_W * NFG + g];
    };
    Face cur, nxt;
    if (nAll > 0) fetch(0, cur);
    for (int j = 0; j < nAll; j++) {
        double nbv[NQ];
#pragma unroll
        for (int k = 0; k < NQ; k++) nbv[k] = f[k * NX + cur.c];
        if (j + 1 < nAll) fetch(j + 1, nxt);
        const int type = cur.type;
        const double Sx = cur.Sx, Sy = cur.Sy, Sz = cur.Sz, w = cur.w;
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            const double nb = nbv[k];
            double ssf;
            if (type == ET_UPPER) ssf = w * (own[k] - nb) + nb;            // P = row, N = col
            else if (type == ET_LOWER) ssf = w * (nb - own[k]) + own[k];   // P = col, N = row
            else if (type == ET_COUPLED) ssf = w * own[k] + (1.0 - w) * nb;
            else ssf = nb;                                                 // patch value
            const double tx = Sx * ssf, ty = Sy * ssf, tz = Sz * ssf;
            if (type == ET_LOWER) { acc[k][0] -= tx; acc[k][1] -= ty; acc[k][2] -= tz; }
            else { acc[k][0] += tx; acc[k][1] += ty; acc[k][2] += tz; }
        }
        cur = nxt;
    }
    const double vol = V[p];
#pragma unroll
    for (int k = 0; k < NQ; k++) {
        grad[(size_t)(k * 3 + 0) * NPH + p] = acc[k][0] / vol;
        grad[(size_t)(k * 3 + 1) * NPH + p] = acc[k][1] / vol;
        grad[(size_t)(k * 3 + 2) * NPH + p] = acc[k][2] / vol;
    }
}

// ------------------------------------------------------------------------------------------------ k_flux
struct DdtPrm {
    int scheme;  // ICSB200_DDT_*
    double rDeltaT, coefft, coefft0, coefft00;
};

struct FluxArgs {
    int NP, NB, F;
    const int *pos2cell, *sliceOff, *rowNAll, *rowNLow, *col, *meta, *gfid;
    const double *geo, *dCoupled, *C, *V, *f, *grad;
    size_t NFG, NX, NPH;
    int limRho, limU, limT;
    SchemePrm sp;
    DdtPrm ddt;
    const double *rdt, *Wold, *Wold2;  // [NP], [5*NP]
    double* src;                       // [5*NPH]
    double* faceFlux;                  // [5*NFG]
    double* recon;                     // [8*NFG] limited U_l U_r E_l E_r per face, reused by the Jacobian kernel
    double* phiB;                      // [NB]
    const double* visc;                // [8*NP] viscous divergences (already divided by V) or null
    const double* mrf;                 // [NFG] MRFFaceVelocity in GPU face order or null (zero field)
};

// reconstruct all NQ scalars of one face.  rowIsOwner: the row cell is the face's owner (P)
__device__ __forceinline__ void reconstructFace(const FluxArgs& a, int p, int c, int type, size_t g, int b, bool needC, FaceState& s)
{
    double L[NQ], R[NQ];
    if (type == ET_PHYS) {
#pragma unroll
        for (int k = 0; k < NQ; k++) L[k] = R[k] = a.f[k * a.NX + c];
    } else {
        const bool rowIsP = type != ET_LOWER;
        const int P = rowIsP ? p : c, N = rowIsP ? c : p;
        V3 d;
        if (type == ET_COUPLED) d = {a.dCoupled[b], a.dCoupled[a.NB + b], a.dCoupled[2 * (size_t)a.NB + b]};
        else d = {a.C[N] - a.C[P], a.C[a.NPH + N] - a.C[a.NPH + P], a.C[2 * a.NPH + N] - a.C[2 * a.NPH + P]};
        const double w = a.geo[G_W * a.NFG + g];
#pragma unroll
        for (int k = 0; k < NQ; k++) {
            if (k == Q_CR && !needC) { L[k] = R[k] = 0.0; continue; }
            const int lim = (k <= Q_P) ? a.limRho : (k <= Q_UZ ? a.limU : a.limT);
            const double phiP = a.f[k * a.NX + P], phiN = a.f[k * a.NX + N];
            V3 gP = {0, 0, 0}, gN = {0, 0, 0};
            if (lim == ICSB200_LIM_VANLEER || lim == ICSB200_LIM_MINMOD) {
                gP = {a.grad[(size_t)(k * 3) * a.NPH + P], a.grad[(size_t)(k * 3 + 1) * a.NPH + P], a.grad[(size_t)(k * 3 + 2) * a.NPH + P]};
                gN = {a.grad[(size_t)(k * 3) * a.NPH + N], a.grad[(size_t)(k * 3 + 1) * a.NPH + N], a.grad[(size_t)(k * 3 + 2) * a.NPH + N]};
            }
            reconLR(lim, type == ET_COUPLED, phiP, phiN, gP, gN, d, w, L[k], R[k]);
        }
    }
    s.rho_l = L[Q_RHO]; s.rho_r = R[Q_RHO]; s.p_l = L[Q_P]; s.p_r = R[Q_P];
    s.U_l = {L[Q_UX], L[Q_UY], L[Q_UZ]}; s.U_r = {R[Q_UX], R[Q_UY], R[Q_UZ]};
    s.c_l = L[Q_CR]; s.c_r = R[Q_CR]; s.E_l = L[Q_E]; s.E_r = R[Q_E]; s.H_l = L[Q_H]; s.H_r = R[Q_H];
}

// Pass 1: every row evaluates the faces it owns (entries at and after rowNLow: upper, coupled and physical faces) ONCE,
// in the face's own orientation, and stores the 5 fluxes in GPU-face order — the row's own entries are consecutive
// GPU face ids per lane, so the stores coalesce like the geometry reads.
template <int SCHEME>
__global__ void __launch_bounds__(128, 4)
k_flux_faces(FluxArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.NP || a.pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)a.sliceOff[p >> 5];
    const int nAll = a.rowNAll[p];
    constexpr bool needC = SCHEME != ICSB200_FLUX_ROE;
    for (int j = a.rowNLow[p]; j < nAll; j++) {
        const size_t e = (base + j) * 32 + lane;
        const int c = a.col[e], m = a.meta[e], type = m & 3;
        const size_t g = a.gfid[e];
        const int b = (m >> 2) - a.F;
        FaceState s;
        reconstructFace(a, p, c, type, g, b, needC, s);
        const V3 Sf = {a.geo[G_SFX * a.NFG + g], a.geo[G_SFY * a.NFG + g], a.geo[G_SFZ * a.NFG + g]};
        const double magSf = a.geo[G_MAGSF * a.NFG + g];
        const double mrf = a.mrf ? a.mrf[g] : 0.0;
        const Flux5 F = faceFlux<SCHEME>(s, Sf, magSf, mrf, a.sp);
        a.faceFlux[g] = F.phi; a.faceFlux[a.NFG + g] = F.phiUp.x; a.faceFlux[2 * a.NFG + g] = F.phiUp.y;
        a.faceFlux[3 * a.NFG + g] = F.phiUp.z; a.faceFlux[4 * a.NFG + g] = F.phiEp;
        if (type == ET_PHYS) a.phiB[b] = F.phi;
        else {
            // the same limited states feed createConvectiveJacobian (convectiveFluxScheme.C:387-400)
            a.recon[g] = s.U_l.x; a.recon[a.NFG + g] = s.U_l.y; a.recon[2 * a.NFG + g] = s.U_l.z;
            a.recon[3 * a.NFG + g] = s.U_r.x; a.recon[4 * a.NFG + g] = s.U_r.y; a.recon[5 * a.NFG + g] = s.U_r.z;
            a.recon[6 * a.NFG + g] = s.E_l; a.recon[7 * a.NFG + g] = s.E_r;
        }
    }
}

// Pass 2: fvc::surfaceIntegrate per row in ascending face id (+ for owned / boundary faces, - where the row is the
// neighbour), then the viscous, dual-time and source terms of residualsUpdate.H.
__global__ void __launch_bounds__(256)
k_flux_gather(FluxArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.NP || a.pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)a.sliceOff[p >> 5];
    double acc[5] = {0, 0, 0, 0, 0};
    const int nAll = a.rowNAll[p], nLow = a.rowNLow[p];
    // all face ids first, then the fluxes of the next face while the current one is added (ascending face id kept)
    double fc[5], fn[5];
    auto fetch = [&](int j, double* o) {
        const size_t g = a.gfid[(base + j) * 32 + lane];
#pragma unroll
        for (int k = 0; k < 5; k++) o[k] = a.faceFlux[k * a.NFG + g];
    };
    if (nAll > 0) fetch(0, fc);
    for (int j = 0; j < nAll; j++) {
        if (j + 1 < nAll) fetch(j + 1, fn);
        if (j < nLow) { acc[0] -= fc[0]; acc[1] -= fc[1]; acc[2] -= fc[2]; acc[3] -= fc[3]; acc[4] -= fc[4]; }
        else { acc[0] += fc[0]; acc[1] += fc[1]; acc[2] += fc[2]; acc[3] += fc[3]; acc[4] += fc[4]; }
#pragma unroll
        for (int k = 0; k < 5; k++) fc[k] = fn[k];
    }
    // residualsUpdate.H: R = -div(phi*) [- (ddt.diag*W - ddt.source)/V]; source = R*V
    const double vol = a.V[p];
    const double rdt = a.rdt[p];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double Rk = -(acc[k] / vol);
        if (a.visc) {  // residualsUpdate.H:25-42: rhoUR += laplacian(muEff,U); += div(tauMC); rhoER += div(sigmaDotU & Sf); += laplacian(alphaEff,e)
            if (k >= 1 && k <= 3) { Rk += a.visc[(size_t)(k - 1) * a.NP + p]; Rk += a.visc[(size_t)(3 + k - 1) * a.NP + p]; }
            else if (k == 4) { Rk += a.visc[(size_t)6 * a.NP + p]; Rk += a.visc[(size_t)7 * a.NP + p]; }
        }
        if (a.ddt.scheme != ICSB200_DDT_STEADY) {
            const double Wk = a.f[(size_t)(Q_W0 + k) * a.NX + p];
            double diag, source;
            if (a.ddt.scheme == ICSB200_DDT_EULER) {
                diag = a.ddt.rDeltaT * vol;
                source = a.ddt.rDeltaT * a.Wold[(size_t)k * a.NP + p] * vol;
            } else {
                diag = (a.ddt.coefft * a.ddt.rDeltaT) * vol;
                source = a.ddt.rDeltaT * vol * (a.ddt.coefft0 * a.Wold[(size_t)k * a.NP + p] - a.ddt.coefft00 * a.Wold2[(size_t)k * a.NP + p]);
            }
            diag += rdt * vol;
            source += rdt * Wk * vol;
            Rk -= (diag * Wk - source) / vol;
        }
        a.src[(size_t)k * a.NPH + p] = Rk * vol;
    }
}


// ------------------------------------------------------------------------------------------------ k_visc
// Viscous part of residualsUpdate.H:16-43 (laminar): per row, in ascending face id, the four surface integrals
//   laplacian(muEff,U), div(tauMC), div(sigmaDotU & Sf), laplacian(alphaEff,eCalc)
// with `Gauss linear corrected` laplacians (snGrad = nonOrthDeltaCoeffs (N - P) + (n - delta nonOrthDeltaCoeffs) &
// interpolate(grad)), tauMC = mu dev2(T(grad U)), patch snGrad of U by patch-field type and the boundary values of
// grad(U) from gaussGrad::correctBoundaryConditions.  Output: the divergences (sum / V) for k_flux to add in order.
struct ViscArgs {
    int NP, NB, F;
    const int *pos2cell, *sliceOff, *rowNAll, *col, *meta, *gfid, *bfPatch;
    const BCDev* bcs;
    const double *geo, *dCoupled, *C, *V, *f, *grad, *gradE, *vic;
    size_t NFG, NX, NPH;
    double mu, alphaEff;
    const double* tr;  // [2][NX] muEff, alphaEff fields (cells, halo and boundary slots) or null: laminar constants
    const int* bfNbrPos;      // [NB] rotational cyclic faces: position of the neighbour patch's face cell
    const int* bfAmiStart;    // [NB+1] AMI stencil of a boundary face (empty: not an AMI face), positions and weights
    const int* amiSrc;
    const double* amiW;
    const double* patchRot;   // [10*nPatches] (rotational flag, forwardT[9])
    double* out;  // [8*NP]
};

__device__ __forceinline__ void gradUAt(const ViscArgs& a, int q, double g[9])  // g[3*i+j] = d_i U_j
{
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) g[3 * i + j] = a.grad[(size_t)((Q_UX + j) * 3 + i) * a.NPH + q];
}
__device__ __forceinline__ void dev2T(const double g[9], double mu, double tau[9])
{
    double A[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) A[3 * i + j] = g[3 * j + i];
    const double tr = A[0] + A[4] + A[8];
    const double sph = (2.0 / 3.0) * tr;
#pragma unroll
    for (int k = 0; k < 9; k++) tau[k] = A[k];
    tau[0] = A[0] - sph; tau[4] = A[4] - sph; tau[8] = A[8] - sph;
#pragma unroll
    for (int k = 0; k < 9; k++) tau[k] = mu * tau[k];
}

__global__ void __launch_bounds__(128, 2)
k_visc(ViscArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.NP || a.pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)a.sliceOff[p >> 5];
    double lap[3] = {0, 0, 0}, dtau[3] = {0, 0, 0}, sg = 0.0, le = 0.0;
    const int nAll = a.rowNAll[p];
    for (int j = 0; j < nAll; j++) {
        const size_t e = (base + j) * 32 + lane;
        const int c = a.col[e], m = a.meta[e], type = m & 3;
        const size_t g = a.gfid[e];
        const int b = (m >> 2) - a.F;
        const double Sf[3] = {a.geo[G_SFX * a.NFG + g], a.geo[G_SFY * a.NFG + g], a.geo[G_SFZ * a.NFG + g]};
        const double magSf = a.geo[G_MAGSF * a.NFG + g];
        double nf[3];
#pragma unroll
        for (int d = 0; d < 3; d++) nf[d] = Sf[d] / magSf;
        double fl[3], ft[3], fs, fe;
        if (type != ET_PHYS) {
            const bool coupled = type == ET_COUPLED;
            const bool rowIsP = type != ET_LOWER;
            const int P = rowIsP ? p : c, N = rowIsP ? c : p;
            const double w = a.geo[G_W * a.NFG + g], dcn = a.geo[G_NONORTH * a.NFG + g];
            double dv[3];
            if (coupled) { dv[0] = a.dCoupled[b]; dv[1] = a.dCoupled[a.NB + b]; dv[2] = a.dCoupled[2 * (size_t)a.NB + b]; }
            else { dv[0] = a.C[N] - a.C[P]; dv[1] = a.C[a.NPH + N] - a.C[a.NPH + P]; dv[2] = a.C[2 * a.NPH + N] - a.C[2 * a.NPH + P]; }
            double corr[3];
#pragma unroll
            for (int d = 0; d < 3; d++) corr[d] = nf[d] - dv[d] * dcn;
            auto lin = [&](double x, double y) { return coupled ? w * x + (1.0 - w) * y : w * (x - y) + y; };
            double gP[9], gN[9], gf[9], tP[9], tN[9], tf[9];
            gradUAt(a, P, gP); gradUAt(a, N, gN);
            const double muP = a.tr ? a.tr[P] : a.mu, muN = a.tr ? a.tr[N] : a.mu;
            const double alP = a.tr ? a.tr[a.NX + P] : a.alphaEff, alN = a.tr ? a.tr[a.NX + N] : a.alphaEff;
            dev2T(gP, muP, tP);
            const double* rot = coupled ? a.patchRot + (size_t)10 * a.bfPatch[b] : nullptr;
            const bool isRot = coupled && rot[0] != 0.0;
            const int amiBeg = coupled ? a.bfAmiStart[b] : 0, amiEnd = coupled ? a.bfAmiStart[b + 1] : 0;
            if (isRot || amiEnd > amiBeg) {
                // cyclicFvPatchField / cyclicAMIFvPatchField<tensor>::patchNeighbourField: transform(forwardT, t) = (T & t) & T.T() on
                // rotational pairs (cyclicFvPatchField.C:130-190).  The halo slot holds transform(forwardT, grad(U_j)) per component,
                // i.e. T & gradU; tauMC is a cell field, so it is the neighbour CELL's tensor (cyclic) or the AMI interpolation of
                // the neighbour cells' tensors (result = 0; result += w_k tau_k) that is taken, then rotated.
                const double* T = rot + 1;
                double h[9], gRaw[9], tr[9], hr[9];
                if (isRot) {
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int j = 0; j < 3; j++) h[3 * i + j] = gN[3 * i] * T[3 * j] + gN[3 * i + 1] * T[3 * j + 1] + gN[3 * i + 2] * T[3 * j + 2];
#pragma unroll
                    for (int k = 0; k < 9; k++) gN[k] = h[k];
                }
                if (amiEnd > amiBeg) {
#pragma unroll
                    for (int k = 0; k < 9; k++) tr[k] = 0.0;
                    for (int q = amiBeg; q < amiEnd; q++) {
                        const int pk = a.amiSrc[q];
                        double tk[9];
                        gradUAt(a, pk, gRaw);
                        dev2T(gRaw, a.tr ? a.tr[pk] : a.mu, tk);
                        const double wk = a.amiW[q];
#pragma unroll
                        for (int k = 0; k < 9; k++) tr[k] += wk * tk[k];
                    }
                } else {
                    gradUAt(a, a.bfNbrPos[b], gRaw);
                    dev2T(gRaw, muN, tr);
                }
                if (!isRot) {
#pragma unroll
                    for (int k = 0; k < 9; k++) tN[k] = tr[k];
                } else {
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int l = 0; l < 3; l++) hr[3 * i + l] = T[3 * i] * tr[l] + T[3 * i + 1] * tr[3 + l] + T[3 * i + 2] * tr[6 + l];
#pragma unroll
                    for (int i = 0; i < 3; i++)
#pragma unroll
                        for (int j = 0; j < 3; j++) tN[3 * i + j] = hr[3 * i] * T[3 * j] + hr[3 * i + 1] * T[3 * j + 1] + hr[3 * i + 2] * T[3 * j + 2];
                }
            } else
                dev2T(gN, muN, tN);
#pragma unroll
            for (int k = 0; k < 9; k++) { gf[k] = lin(gP[k], gN[k]); tf[k] = lin(tP[k], tN[k]); }
            const double muf = lin(muP, muN), alf = lin(alP, alN);
            double UP[3], UN[3], Uf[3];
#pragma unroll
            for (int d = 0; d < 3; d++) { UP[d] = a.f[(size_t)(Q_UX + d) * a.NX + P]; UN[d] = a.f[(size_t)(Q_UX + d) * a.NX + N]; Uf[d] = lin(UP[d], UN[d]); }
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const double snG = dcn * (UN[d] - UP[d]) + (corr[0] * gf[d] + corr[1] * gf[3 + d] + corr[2] * gf[6 + d]);
                fl[d] = muf * snG * magSf;
                ft[d] = Sf[0] * tf[d] + Sf[1] * tf[3 + d] + Sf[2] * tf[6 + d];
            }
            double sd[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double a0 = muf * gf[3 * i] + tf[3 * i], a1 = muf * gf[3 * i + 1] + tf[3 * i + 1], a2 = muf * gf[3 * i + 2] + tf[3 * i + 2];
                sd[i] = a0 * Uf[0] + a1 * Uf[1] + a2 * Uf[2];
            }
            fs = sd[0] * Sf[0] + sd[1] * Sf[1] + sd[2] * Sf[2];
            double gEf[3];
#pragma unroll
            for (int d = 0; d < 3; d++) gEf[d] = lin(a.gradE[(size_t)d * a.NPH + P], a.gradE[(size_t)d * a.NPH + N]);
            const double eP = a.f[(size_t)Q_EC * a.NX + P], eN = a.f[(size_t)Q_EC * a.NX + N];
            const double snE = dcn * (eN - eP) + (corr[0] * gEf[0] + corr[1] * gEf[1] + corr[2] * gEf[2]);
            fe = alf * snE * magSf;
        } else {
            // physical patch face: P = row, patch values at slot c
            const double dc = a.geo[G_DELTA * a.NFG + g];
            const int pi = a.bfPatch[b];
            const BCDev& bc = a.bcs[pi];
            double Ui[3], Ub[3], sn[3];
#pragma unroll
            for (int d = 0; d < 3; d++) { Ui[d] = a.f[(size_t)(Q_UX + d) * a.NX + p]; Ub[d] = a.f[(size_t)(Q_UX + d) * a.NX + c]; }
            const int kind = bc.kind[ICSB200_FIELD_U];
            if (kind == ICSB200_BC_FIXEDVALUE || kind == ICSB200_BC_PRESSUREINLETOUTLETVELOCITY) {
#pragma unroll
                for (int d = 0; d < 3; d++) sn[d] = dc * (Ub[d] - Ui[d]);
            } else if (kind == ICSB200_BC_SLIP) {
                const double xx = 1.0 - 2.0 * (nf[0] * nf[0]), xy = 0.0 - 2.0 * (nf[0] * nf[1]), xz = 0.0 - 2.0 * (nf[0] * nf[2]);
                const double yy = 1.0 - 2.0 * (nf[1] * nf[1]), yz = 0.0 - 2.0 * (nf[1] * nf[2]), zz = 1.0 - 2.0 * (nf[2] * nf[2]);
                const double t[3] = {xx * Ui[0] + xy * Ui[1] + xz * Ui[2], xy * Ui[0] + yy * Ui[1] + yz * Ui[2], xz * Ui[0] + yz * Ui[1] + zz * Ui[2]};
#pragma unroll
                for (int d = 0; d < 3; d++) sn[d] = (t[d] - Ui[d]) * (dc / 2.0);
            } else if (kind == ICSB200_BC_INLETOUTLET) {
                const double vfrac = 1.0 - a.vic[a.NB + b];
#pragma unroll
                for (int d = 0; d < 3; d++) sn[d] = vfrac * (bc.P(ICSB200_FIELD_U, b)[d] - Ui[d]) * dc + (1.0 - vfrac) * 0.0;
            } else {
                sn[0] = sn[1] = sn[2] = 0.0;
            }
            double gP[9], gb[9], tb[9];
            gradUAt(a, p, gP);
#pragma unroll
            for (int jj = 0; jj < 3; jj++) {
                const double ng = nf[0] * gP[jj] + nf[1] * gP[3 + jj] + nf[2] * gP[6 + jj];
#pragma unroll
                for (int i = 0; i < 3; i++) gb[3 * i + jj] = gP[3 * i + jj] + nf[i] * (sn[jj] - ng);
            }
            const double muB = a.tr ? a.tr[c] : a.mu, alB = a.tr ? a.tr[a.NX + c] : a.alphaEff;  // patch values of muEff / alphaEff
            dev2T(gb, muB, tb);
#pragma unroll
            for (int d = 0; d < 3; d++) {
                fl[d] = muB * sn[d] * magSf;
                ft[d] = Sf[0] * tb[d] + Sf[1] * tb[3 + d] + Sf[2] * tb[6 + d];
            }
            double sd[3];
#pragma unroll
            for (int i = 0; i < 3; i++) {
                const double a0 = muB * gb[3 * i] + tb[3 * i], a1 = muB * gb[3 * i + 1] + tb[3 * i + 1], a2 = muB * gb[3 * i + 2] + tb[3 * i + 2];
                sd[i] = a0 * Ub[0] + a1 * Ub[1] + a2 * Ub[2];
            }
            fs = sd[0] * Sf[0] + sd[1] * Sf[1] + sd[2] * Sf[2];
            const double eP = a.f[(size_t)Q_EC * a.NX + p], eB = a.f[(size_t)Q_EC * a.NX + c];
            fe = alB * (dc * (eB - eP)) * magSf;
        }
        if (type == ET_LOWER) {
            lap[0] -= fl[0]; lap[1] -= fl[1]; lap[2] -= fl[2]; dtau[0] -= ft[0]; dtau[1] -= ft[1]; dtau[2] -= ft[2]; sg -= fs; le -= fe;
        } else {
            lap[0] += fl[0]; lap[1] += fl[1]; lap[2] += fl[2]; dtau[0] += ft[0]; dtau[1] += ft[1]; dtau[2] += ft[2]; sg += fs; le += fe;
        }
    }
    const double vol = a.V[p];
#pragma unroll
    for (int d = 0; d < 3; d++) { a.out[(size_t)d * a.NP + p] = lap[d] / vol; a.out[(size_t)(3 + d) * a.NP + p] = dtau[d] / vol; }
    a.out[(size_t)6 * a.NP + p] = sg / vol;
    a.out[(size_t)7 * a.NP + p] = le / vol;
}

}  // namespace

DdtPrm makeDdt(const icsb200_ctx* c)
{
    DdtPrm d{};
    d.scheme = c->sch.ddt_scheme;
    if (d.scheme == ICSB200_DDT_EULER) d.rDeltaT = 1.0 / c->sch.delta_t;
    else if (d.scheme == ICSB200_DDT_BACKWARD) {
        double deltaT = c->sch.delta_t;
        d.rDeltaT = 1.0 / deltaT;
        double deltaT0 = (c->timeIndex < 2) ? ICS_GREAT : deltaT;  // backwardDdtScheme::deltaT0_(vf)
        d.coefft = 1 + deltaT / (deltaT + deltaT0);
        d.coefft00 = deltaT * deltaT / (deltaT0 * (deltaT + deltaT0));
        d.coefft0 = d.coefft + d.coefft00;
    }
    return d;
}

int ics_gradients(icsb200_ctx* c)
{
    {
        LaunchScope ls(c, TM_GRAD);
        // two passes of 4 scalars: half the registers, twice the resident warps to hide the neighbour gathers
        static_assert(NQ == 8, "k_grad passes assume 8 reconstructed scalars");
        for (int half = 0; half < 2; half++)
            k_grad<4><<<gridFor(c->NP, 128), 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_meta, c->d_gfid,
                                                                 c->d_geo, c->NFG, c->d_V, c->d_fields + (size_t)half * 4 * c->NX, c->NX,
                                                                 c->d_grad + (size_t)half * 12 * c->NPH, c->NPH);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    // every gradient is a vector: on rotational cyclic patches each is rotated on its own (transform(forwardT, grad(phi)), also
    // for phi = "U.component(i)": cyclicFvPatchField<vector> of a field whose name does not start with "U")
    unsigned gradMask = 0;
    for (int k = 0; k < NQ; k++) gradMask |= 1u << (3 * k);
    int r = ics_halo_fields(c, c->d_grad, c->NPH, NQ * 3, gradMask);
    if (r || !(c->mu > 0)) return r;
    // viscous runs: gradient of eCalc for the non-orthogonal correction of laplacian(alphaEff, e)
    if (!c->d_gradE) {
        if ((r = devAlloc(c, &c->d_gradE, (size_t)3 * c->NPH))) return r;
        CUDA_TRY(c, cudaMemsetAsync(c->d_gradE, 0, sizeof(double) * 3 * c->NPH, c->stream));
    }
    {
        LaunchScope ls(c, TM_GRAD);
        k_grad<1><<<gridFor(c->NP, 128), 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_meta, c->d_gfid,
                                                             c->d_geo, c->NFG, c->d_V, c->q(Q_EC), c->NX, c->d_gradE, c->NPH);
    }
    CUDA_TRY(c, cudaGetLastError());
    return ics_halo_fields(c, c->d_gradE, c->NPH, 3, 1u);
}

int ics_flux_residual(icsb200_ctx* c, bool storeFaceFlux)
{
    (void)storeFaceFlux;  // the face fluxes are always stored now: each face is evaluated once
    if (!c->d_faceFlux) {
        int r = devAlloc(c, &c->d_faceFlux, (size_t)5 * c->NFG);
        if (r) return r;
    }
    if (!c->d_faceRecon) {
        int r = devAlloc(c, &c->d_faceRecon, (size_t)8 * c->NFG);
        if (r) return r;
    }
    FluxArgs a{};
    a.NP = c->NP; a.NB = c->NB; a.F = c->F;
    a.pos2cell = c->d_pos2cell; a.sliceOff = c->d_sliceOff; a.rowNAll = c->d_rowNAll; a.rowNLow = c->d_rowNLow; a.col = c->d_col; a.meta = c->d_meta; a.gfid = c->d_gfid;
    a.geo = c->d_geo; a.dCoupled = c->d_dCoupled; a.C = c->d_C; a.V = c->d_V; a.f = c->d_fields; a.grad = c->d_grad;
    a.NFG = c->NFG; a.NX = c->NX; a.NPH = c->NPH;
    a.limRho = c->sch.limiter_rho; a.limU = c->sch.limiter_U; a.limT = c->sch.limiter_T;
    a.sp = SchemePrm{c->gamma, c->sch.entropy_fix_coeff, c->sch.low_mach_ausm};
    a.ddt = makeDdt(c);
    a.rdt = c->d_rdt; a.Wold = c->d_Wold; a.Wold2 = c->d_Wold2;
    a.src = c->d_src;
    a.faceFlux = c->d_faceFlux;
    a.recon = c->d_faceRecon;
    a.phiB = c->d_phiB;
    a.mrf = c->d_mrfFace;
    a.visc = nullptr;
    if (c->mu > 0) {  // if (!inviscid)  (createFields.H:37-45)
        if (!c->d_visc) { int r = devAlloc(c, &c->d_visc, (size_t)8 * c->NP); if (r) return r; }
        ViscArgs v{};
        v.NP = c->NP; v.NB = c->NB; v.F = c->F;
        v.pos2cell = c->d_pos2cell; v.sliceOff = c->d_sliceOff; v.rowNAll = c->d_rowNAll; v.col = c->d_col; v.meta = c->d_meta; v.gfid = c->d_gfid;
        v.bfPatch = c->d_bfPatch; v.bcs = c->d_bc;
        v.geo = c->d_geo; v.dCoupled = c->d_dCoupled; v.C = c->d_C; v.V = c->d_V; v.f = c->d_fields; v.grad = c->d_grad; v.gradE = c->d_gradE; v.vic = c->d_vic;
        v.NFG = c->NFG; v.NX = c->NX; v.NPH = c->NPH;
        v.mu = c->mu; v.alphaEff = c->gamma * (c->mu / c->Pr);
        v.tr = c->d_transport;
        v.bfNbrPos = c->d_bfNbrPos; v.patchRot = c->d_patchRot;
        v.bfAmiStart = c->d_bfAmiStart; v.amiSrc = c->d_amiAllSrc; v.amiW = c->d_amiAllW;
        v.out = c->d_visc;
        LaunchScope ls(c, TM_FLUX);
        k_visc<<<gridFor(c->NP, 128), 128, 0, c->stream>>>(v);
        CUDA_TRY(c, cudaGetLastError());
        a.visc = c->d_visc;
    }
    {
        LaunchScope ls(c, TM_FLUX);
        const int grid = gridFor(c->NP, 128);
        if (c->sch.flux_scheme == ICSB200_FLUX_HLLC) k_flux_faces<ICSB200_FLUX_HLLC><<<grid, 128, 0, c->stream>>>(a);
        else if (c->sch.flux_scheme == ICSB200_FLUX_ROE) k_flux_faces<ICSB200_FLUX_ROE><<<grid, 128, 0, c->stream>>>(a);
        else if (c->sch.flux_scheme == ICSB200_FLUX_RUSANOV) k_flux_faces<ICSB200_FLUX_RUSANOV><<<grid, 128, 0, c->stream>>>(a);
        else k_flux_faces<ICSB200_FLUX_AUSMPLUSUP><<<grid, 128, 0, c->stream>>>(a);
        k_flux_gather<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(a);
        c->launches++;
    }
    CUDA_TRY(c, cudaGetLastError());
    c->fluxValid = true;
    c->reconValid = true;
    c->srcMrfApplied = false;
    return ics_hb_source(c);  // Harmonic Balance: sources += -V sum_K D[J][K] W_K
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int icsb200_calc_flux(icsb200_ctx* c, double* phi, double* phiUp, double* phiEp)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "calc_flux: state not set");
    cudaSetDevice(c->device);
    int r;
    if ((r = ics_gradients(c))) return r;
    const bool want = phi || phiUp || phiEp;
    if ((r = ics_flux_residual(c, want))) return r;
    if (want) {
        std::vector<double> h((size_t)5 * c->NFG);
        CUDA_TRY(c, cudaMemcpyAsync(h.data(), c->d_faceFlux, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        const size_t NFG = c->NFG;
        if (phi) std::fill(phi, phi + c->FT, 0.0);
        if (phiUp) std::fill(phiUp, phiUp + 3 * (size_t)c->FT, 0.0);
        if (phiEp) std::fill(phiEp, phiEp + c->FT, 0.0);
        for (size_t g = 0; g < NFG; g++) {
            const size_t f = c->h_gf2ref[g];
            if (phi) phi[f] = h[g];
            if (phiUp) { phiUp[3 * f] = h[NFG + g]; phiUp[3 * f + 1] = h[2 * NFG + g]; phiUp[3 * f + 2] = h[3 * NFG + g]; }
            if (phiEp) phiEp[f] = h[4 * NFG + g];
        }
    } else {
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    return 0;
}

extern "C" int icsb200_residual(icsb200_ctx* c, double* rhoR, double* rhoUR, double* rhoER)
{
    if (!c->fluxValid) return ics_fail(c, ICSB200_ESTATE, "residual: call calc_flux first");
    int r = 0;
    if (rhoR && (r = ics_download_cells(c, rhoR, 1, c->d_src, c->NPH))) return r;
    if (rhoUR && (r = ics_download_cells(c, rhoUR, 3, c->d_src + c->NPH, c->NPH))) return r;
    if (rhoER && (r = ics_download_cells(c, rhoER, 1, c->d_src + 4 * (size_t)c->NPH, c->NPH))) return r;
    return 0;
}
