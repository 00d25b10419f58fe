// oracle_flux.cpp — convective flux schemes, residual, pseudo time step, Jacobian assembly, field update.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
//
// Follows, expression by expression (same operand order, no FMA contraction — build with -ffp-contract=off):
//   hllcFluxScheme.C:70-240, roeFluxScheme.C:41-241 + 276-409, ausmPlusUpFluxScheme.C:73-299,
//   cellFaceFunctions.H:490-660, residualsUpdate.H:1-83, setCoAndDeltaT.H:1-173, outerLoop.H:61-64,
//   dualTimeDdtScheme.C:111-126, convectiveFluxScheme.C:47-120 + 219-546, blockFvMatrix.C:211-326,
//   viscousFluxScheme.C:220-246, blockFvOperatorsTemplates.C:499-557, updateFields.H:1-104,
//   boundLocalTimeStep.H:1-98.
// The reference evaluates whole-field expressions; this file evaluates the same expressions face by face.
#include "oracle_internal.hpp"

namespace orc {

namespace {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double operator&(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 operator^(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double magSqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
struct T9 { double v[9]; };  // xx xy xz yx yy yz zx zy zz
inline T9 outer(V3 a, V3 b) { return {{a.x * b.x, a.x * b.y, a.x * b.z, a.y * b.x, a.y * b.y, a.y * b.z, a.z * b.x, a.z * b.y, a.z * b.z}}; }
inline T9 operator*(T9 a, double s) { T9 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] * s; return r; }
inline T9 operator*(double s, T9 a) { T9 r; for (int i = 0; i < 9; i++) r.v[i] = s * a.v[i]; return r; }
inline T9 operator+(T9 a, T9 b) { T9 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
inline T9 operator-(T9 a, T9 b) { T9 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
inline V3 operator&(T9 t, V3 v) { return {t.v[0] * v.x + t.v[1] * v.y + t.v[2] * v.z, t.v[3] * v.x + t.v[4] * v.y + t.v[5] * v.z, t.v[6] * v.x + t.v[7] * v.y + t.v[8] * v.z}; }
inline V3 operator&(V3 v, T9 t) { return {v.x * t.v[0] + v.y * t.v[3] + v.z * t.v[6], v.x * t.v[1] + v.y * t.v[4] + v.z * t.v[7], v.x * t.v[2] + v.y * t.v[5] + v.z * t.v[8]}; }
inline T9 operator&(T9 a, T9 b)
{
    T9 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.v[3 * i + j] = a.v[3 * i] * b.v[j] + a.v[3 * i + 1] * b.v[3 + j] + a.v[3 * i + 2] * b.v[6 + j];
    return r;
}
const T9 I9 = {{1, 0, 0, 0, 1, 0, 0, 0, 1}};

// reconstructed left/right face states + geometry of one face
struct FaceLR {
    double rho_l, rho_r, p_l, p_r, c_l, c_r, E_l, E_r, H_l, H_r;
    V3 U_l, U_r;
};

// derived vol fields (cells + boundary slots) the schemes reconstruct
struct Derived {
    vecd Ux, Uy, Uz, c, E, H;
};

void derivedFields(Ctx& c, Derived& d, int cKind /*0: max(sqrt(gamma/psi),VSMALL)  1: sqrt(gamma/psi)  2: critical (AUSM)*/)
{
    const Mesh& m = c.m;
    size_t n = (size_t)m.N + m.NB;
    d.Ux.resize(n); d.Uy.resize(n); d.Uz.resize(n); d.c.resize(n); d.E.resize(n); d.H.resize(n);
    for (size_t i = 0; i < n; i++) {
        d.Ux[i] = c.U[3 * i]; d.Uy[i] = c.U[3 * i + 1]; d.Uz[i] = c.U[3 * i + 2];
        double he = c.Cv * c.T[i];                                            // thermo.he(p, T)
        d.E[i] = he + 0.5 * (d.Ux[i] * d.Ux[i] + d.Uy[i] * d.Uy[i] + d.Uz[i] * d.Uz[i]);
        d.H[i] = std::max(d.E[i], SMALL) + std::max(c.p[i] / c.rho[i], SMALL);
        if (cKind == 2) d.c[i] = std::sqrt(2.0 * (c.gamma - 1.0) / (c.gamma + 1.0) * d.H[i]);
        else {
            d.c[i] = std::sqrt(c.gamma / c.psi[i]);
            if (cKind == 0) d.c[i] = std::max(d.c[i], VSMALL);
        }
    }
    // empty-patch slots hold zeros from initialisation; keep them finite
    for (auto& p : m.patches)
        if (m.empty(p))
            for (int f = p.start; f < p.start + p.size; f++) { size_t s = m.N + f - m.F; d.c[s] = d.E[s] = d.H[s] = 0; }
    // coupled patches: patchNeighbourField of an expression field is taken from ITS internal field (cyclic / processor: the
    // neighbour cell's value, identical to evaluating the expression on the copied primitives; cyclicAMI: the
    // AMI-interpolated cell values, which is not the expression of the interpolated primitives)
    syncCoupled(c, d.c, 1); applyPhaseLag(c, d.c, 1, cKind == 2 ? LAG_C2 : (cKind == 1 ? LAG_C1 : LAG_C0));
    syncCoupled(c, d.E, 1); applyPhaseLag(c, d.E, 1, LAG_E);
    syncCoupled(c, d.H, 1); applyPhaseLag(c, d.H, 1, LAG_H);
}

struct Recon {
    vecd rho_l, rho_r, p_l, p_r, Ux_l, Ux_r, Uy_l, Uy_r, Uz_l, Uz_r, c_l, c_r, E_l, E_r, H_l, H_r;
};

void reconstructAll(Ctx& c, const Derived& d, Recon& r, bool needC)
{
    interpolateLimitedLR(c, c.rho, c.sch.limiter_rho, r.rho_l, r.rho_r);
    interpolateLimitedLR(c, c.p, c.sch.limiter_rho, r.p_l, r.p_r);
    interpolateLimitedLR(c, d.Ux, c.sch.limiter_U, r.Ux_l, r.Ux_r);
    interpolateLimitedLR(c, d.Uy, c.sch.limiter_U, r.Uy_l, r.Uy_r);
    interpolateLimitedLR(c, d.Uz, c.sch.limiter_U, r.Uz_l, r.Uz_r);
    if (needC) interpolateLimitedLR(c, d.c, c.sch.limiter_T, r.c_l, r.c_r);
    interpolateLimitedLR(c, d.E, c.sch.limiter_T, r.E_l, r.E_r);
    interpolateLimitedLR(c, d.H, c.sch.limiter_T, r.H_l, r.H_r);
}

inline bool faceActive(const Mesh& m, int f, const std::vector<char>& emptyFace) { return f < m.F || !emptyFace[f - m.F]; }

std::vector<char> emptyMask(const Mesh& m)
{
    std::vector<char> e(m.NB, 0);
    for (auto& p : m.patches)
        if (m.empty(p)) for (int f = p.start; f < p.start + p.size; f++) e[f - m.F] = 1;
    return e;
}

// ---------------------------------------------------------------------------------------------- HLLC
void fluxHLLC(const Ctx& c, const FaceLR& s, V3 Sf, double magSf, double mrf, double& phi, V3& phiUp, double& phiEp)
{
    V3 n = Sf / magSf;
    double coefR = std::sqrt(std::max(VSMALL, s.rho_r) / std::max(VSMALL, s.rho_l));
    V3 uAvg = (coefR * s.U_r + s.U_l) / (coefR + 1.0);
    double HAvg = (coefR * s.H_r + s.H_l) / (coefR + 1.0);
    double gammaf = c.gamma;  // fvc::interpolate(gamma) of a uniform field
    double cAvg = std::sqrt(std::fabs((gammaf - 1.0) * (HAvg - 0.5 * magSqr(uAvg))));
    double uMag_l = s.U_l & n, uMag_r = s.U_r & n, uMagAvg = uAvg & n;
    uMagAvg -= mrf; uMag_l -= mrf; uMag_r -= mrf;
    double Sl = std::min(uMag_l - s.c_l, uMagAvg - cAvg);
    double Sr = std::max(uMag_r + s.c_r, uMagAvg + cAvg);
    double Sm = (s.rho_r * uMag_r * (Sr - uMag_r) - s.rho_l * uMag_l * (Sl - uMag_l) + s.p_l - s.p_r) /
                (s.rho_r * (Sr - uMag_r) - s.rho_l * (Sl - uMag_l));
    double coefSl = pos0(Sl), coefSr = neg(Sr), coefSm = pos0(Sm);
    double coefSlm = (1.0 - coefSl) * coefSm;
    double coefSmr = (1.0 - coefSm) * (1.0 - coefSr);
    double fluxRhoStar_l = Sm / (Sl - Sm) * ((Sl - uMag_l) * s.rho_l);
    double fluxRhoStar_r = Sm / (Sr - Sm) * ((Sr - uMag_r) * s.rho_r);
    phi = (coefSl * s.rho_l * uMag_l + coefSlm * fluxRhoStar_l + coefSmr * fluxRhoStar_r + coefSr * s.rho_r * uMag_r) * magSf;
    double pStar_l = s.rho_l * (uMag_l - Sl) * (uMag_l - Sm) + s.p_l;
    double pStar_r = s.rho_r * (uMag_r - Sr) * (uMag_r - Sm) + s.p_r;
    V3 rhoUStar_l = 1.0 / (Sl - Sm) * ((Sl - uMag_l) * s.rho_l * s.U_l + (pStar_l - s.p_l) * n);
    V3 rhoUStar_r = 1.0 / (Sr - Sm) * ((Sr - uMag_r) * s.rho_r * s.U_r + (pStar_r - s.p_r) * n);
    V3 fluxRhoUStar_l = Sm * rhoUStar_l + pStar_l * n;
    V3 fluxRhoUStar_r = Sm * rhoUStar_r + pStar_r * n;
    phiUp = (coefSl * (s.rho_l * uMag_l * s.U_l + s.p_l * n) + coefSlm * fluxRhoUStar_l + coefSmr * fluxRhoUStar_r +
             coefSr * (s.rho_r * uMag_r * s.U_r + s.p_r * n)) * magSf;
    double rhoEStar_l = 1.0 / (Sl - Sm) * ((Sl - uMag_l) * (s.rho_l * s.E_l) - s.p_l * uMag_l + pStar_l * Sm);
    double rhoEStar_r = 1.0 / (Sr - Sm) * ((Sr - uMag_r) * (s.rho_r * s.E_r) - s.p_r * uMag_r + pStar_r * Sm);
    double fluxRhoEStar_l = Sm * (rhoEStar_l + pStar_l) + pStar_l * mrf;
    double fluxRhoEStar_r = Sm * (rhoEStar_r + pStar_r) + pStar_r * mrf;
    phiEp = (coefSl * s.rho_l * s.H_l * uMag_l + coefSlm * fluxRhoEStar_l + coefSmr * fluxRhoEStar_r +
             coefSr * s.rho_r * s.H_r * uMag_r) * magSf;
}

// ---------------------------------------------------------------------------------------------- ROE
void fluxROE(const Ctx& c, const FaceLR& s, V3 Sf, double magSf, double mrf, double& phi, V3& phiUp, double& phiEp)
{
    V3 n = Sf / magSf;
    double coefR = std::sqrt(std::max(VSMALL, s.rho_r) / std::max(VSMALL, s.rho_l));
    double RoeDensity = coefR * s.rho_l;
    V3 RoeVelocity = (coefR * s.U_r + s.U_l) / (coefR + 1.0);
    double RoeEnthalpy = (coefR * s.H_r + s.H_l) / (coefR + 1.0);
    double gammaInterp = c.gamma;
    double RoeSoundSpeed = std::sqrt(std::fabs((gammaInterp - 1.0) * (RoeEnthalpy - 0.5 * magSqr(RoeVelocity))));
    double uMag_l = s.U_l & n, uMag_r = s.U_r & n;
    double uProjRoe = RoeVelocity & n;
    V3 rhoU_l = s.rho_l * s.U_l, rhoU_r = s.rho_r * s.U_r;
    double rhoE_l = s.rho_l * s.E_l, rhoE_r = s.rho_r * s.E_r;
    uProjRoe -= mrf;
    // getRoeDissipation (roeFluxScheme.C:41-241)
    double a1 = gammaInterp - 1;
    double theta = 0.5 * a1 * magSqr(RoeVelocity);
    double c2 = sqr(RoeSoundSpeed);
    double a2 = 1 / (RoeDensity * RoeSoundSpeed * std::sqrt(2.0));
    double a3 = RoeDensity / (RoeSoundSpeed * std::sqrt(2.0));
    double a4 = (theta + c2) / a1;
    double a5 = 1 - theta / c2;
    double a6 = theta / a1;
    V3 invP11 = (n * a5) - (RoeVelocity ^ n) / RoeDensity;
    T9 invP12 = outer(a1 / c2 * n, RoeVelocity);
    invP12.v[1] += n.z / RoeDensity;
    invP12.v[2] -= n.y / RoeDensity;
    invP12.v[3] -= n.z / RoeDensity;
    invP12.v[5] += n.x / RoeDensity;
    invP12.v[6] += n.y / RoeDensity;
    invP12.v[7] -= n.x / RoeDensity;
    V3 invP13 = -a1 / c2 * n;
    double invP21 = a2 * (theta - RoeSoundSpeed * uProjRoe);
    V3 invP22 = -a2 * (a1 * RoeVelocity - RoeSoundSpeed * n);
    double invP23 = a1 * a2;
    double invP31 = a2 * (theta + RoeSoundSpeed * uProjRoe);
    V3 invP32 = -a2 * (a1 * RoeVelocity + RoeSoundSpeed * n);
    double invP33 = a1 * a2;
    double Lambda1 = std::fabs(uProjRoe);
    double Lambda2 = std::fabs(uProjRoe + RoeSoundSpeed);
    double Lambda3 = std::fabs(uProjRoe - RoeSoundSpeed);
    double epsilon = c.sch.entropy_fix_coeff * std::max(Lambda2, Lambda3);
    if (Lambda1 < epsilon) Lambda1 = (sqr(Lambda1) + sqr(epsilon)) / (2.0 * epsilon);
    if (Lambda2 < epsilon) Lambda2 = (sqr(Lambda2) + sqr(epsilon)) / (2.0 * epsilon);
    if (Lambda3 < epsilon) Lambda3 = (sqr(Lambda3) + sqr(epsilon)) / (2.0 * epsilon);
    V3 P11 = n;
    double P12 = a3, P13 = a3;
    T9 P21 = outer(RoeVelocity, n);
    P21.v[1] -= n.z * RoeDensity;
    P21.v[2] += n.y * RoeDensity;
    P21.v[3] += n.z * RoeDensity;
    P21.v[5] -= n.x * RoeDensity;
    P21.v[6] -= n.y * RoeDensity;
    P21.v[7] += n.x * RoeDensity;
    V3 P22 = a3 * (RoeVelocity + RoeSoundSpeed * n);
    V3 P23 = a3 * (RoeVelocity - RoeSoundSpeed * n);
    V3 P31 = n * a6 + RoeDensity * (RoeVelocity ^ n);
    double P32 = a3 * (a4 + RoeSoundSpeed * uProjRoe);
    double P33 = a3 * (a4 - RoeSoundSpeed * uProjRoe);
    double dissContByRho = ((P11 * Lambda1) & invP11) + (Lambda2 * P12 * invP21) + (Lambda3 * P13 * invP31);
    V3 dissContByRhoU = ((P11 * Lambda1) & invP12) + (Lambda2 * P12 * invP22) + (Lambda3 * P13 * invP32);
    double dissContByRhoE = ((P11 * Lambda1) & invP13) + (Lambda2 * P12 * invP23) + (Lambda3 * P13 * invP33);
    V3 dissMomByRho = ((P21 * Lambda1) & invP11) + (Lambda2 * P22 * invP21) + (Lambda3 * P23 * invP31);
    T9 dissMomByRhoU = ((P21 * Lambda1) & invP12) + outer(Lambda2 * P22, invP22) + outer(Lambda3 * P23, invP32);
    V3 dissMomByRhoE = ((P21 * Lambda1) & invP13) + (Lambda2 * P22 * invP23) + (Lambda3 * P23 * invP33);
    double dissEnergyByRho = ((P31 * Lambda1) & invP11) + (Lambda2 * P32 * invP21) + (Lambda3 * P33 * invP31);
    V3 dissEnergyByRhoU = ((P31 * Lambda1) & invP12) + (Lambda2 * P32 * invP22) + (Lambda3 * P33 * invP32);
    double dissEnergyByRhoE = ((P31 * Lambda1) & invP13) + (Lambda2 * P32 * invP23) + (Lambda3 * P33 * invP33);
    // roeFluxScheme.C:376-408
    double diffRho = s.rho_r - s.rho_l;
    V3 diffRhoU = rhoU_r - rhoU_l;
    double diffRhoE = rhoE_r - rhoE_l;
    phi = -0.5 * magSf * (dissContByRho * diffRho + (dissContByRhoU & diffRhoU) + dissContByRhoE * diffRhoE);
    phiUp = (-0.5 * magSf) * (dissMomByRho * diffRho + (dissMomByRhoU & diffRhoU) + dissMomByRhoE * diffRhoE);
    phiEp = -0.5 * magSf * (dissEnergyByRho * diffRho + (dissEnergyByRhoU & diffRhoU) + dissEnergyByRhoE * diffRhoE);
    double rhoUNorm_l = s.rho_l * uMag_l, rhoUNorm_r = s.rho_r * uMag_r;
    phi += 0.5 * magSf * (rhoUNorm_l + rhoUNorm_r);
    phiUp = phiUp + (0.5 * magSf) * (rhoUNorm_l * s.U_l + rhoUNorm_r * s.U_r + n * (s.p_l + s.p_r));
    phiEp += 0.5 * magSf * (rhoUNorm_l * s.H_l + rhoUNorm_r * s.H_r);
    phi -= 0.5 * magSf * mrf * (s.rho_l + s.rho_r);
    phiUp = phiUp - (0.5 * magSf * mrf) * (rhoU_l + rhoU_r);
    phiEp -= 0.5 * magSf * mrf * (rhoE_l + rhoE_r);
}

// ---------------------------------------------------------------------------------------------- AUSM+up
void fluxAUSM(const Ctx& c, const FaceLR& s, V3 Sf, double magSf, double mrf, double& phi, V3& phiUp, double& phiEp)
{
    double phi_L = s.U_l & Sf, phi_R = s.U_r & Sf;
    double un_L = phi_L / magSf, un_R = phi_R / magSf;
    un_L -= mrf; un_R -= mrf;
    double c_L = sqr(s.c_l) / std::max(s.c_l, un_L);
    double c_R = sqr(s.c_r) / std::max(s.c_r, -un_R);
    double c_face = std::min(c_L, c_R);
    double Mach_L = un_L / c_face;
    double Mach_plus_L, p_plus_L;
    if (std::fabs(Mach_L) < 1.0) {
        double ML2p = 0.25 * sqr(Mach_L + 1), ML2m = -0.25 * sqr(Mach_L - 1);
        Mach_plus_L = ML2p * (1 - 2 * ML2m);
        p_plus_L = ML2p * (2 - Mach_L - 3 * Mach_L * ML2m);
    } else {
        Mach_plus_L = std::max(Mach_L, 0.0);
        p_plus_L = (Mach_L > 0 ? 1.0 : 0.0);
    }
    double Mach_R = un_R / c_face;
    double Mach_minus_R, p_minus_R;
    if (std::fabs(Mach_R) < 1.0) {
        double MR2m = -0.25 * sqr(Mach_R - 1), MR2p = 0.25 * sqr(Mach_R + 1);
        Mach_minus_R = MR2m * (1 + 2 * MR2p);
        p_minus_R = MR2m * (-2 - Mach_R + 3 * Mach_R * MR2p);
    } else {
        Mach_minus_R = std::min(Mach_R, 0.0);
        p_minus_R = (Mach_R < 0 ? 1.0 : 0.0);
    }
    double Mach_1_2 = Mach_plus_L + Mach_minus_R;
    double p_1_2 = p_plus_L * s.p_l + p_minus_R * s.p_r;
    double M_mean = 0.5 * (sqr(un_L) + sqr(un_R)) / sqr(c_face);
    double MDiff = -0.25 * std::max((1.0 - M_mean), 0.0) * (s.p_r - s.p_l) / (0.5 * (s.rho_l + s.rho_r) * sqr(c_face));
    if ((Mach_1_2 > 0.0 && Mach_1_2 + MDiff <= 0.0) || (Mach_1_2 < 0.0 && Mach_1_2 + MDiff >= 0.0)) Mach_1_2 += 0.2 * MDiff;
    else Mach_1_2 += MDiff;
    if (c.sch.low_mach_ausm) {
        double pDiff = -0.25 * p_plus_L * p_minus_R * (s.rho_l + s.rho_r) * c_face * (un_R - un_L);
        p_1_2 += pDiff;
    }
    bool left = Mach_1_2 >= 0;
    V3 U_f = left ? s.U_l : s.U_r;
    double rhoa_LR = Mach_1_2 * c_face * (left ? s.rho_l : s.rho_r);
    V3 rhoaU_LR = rhoa_LR * U_f;
    double rhoah_LR = rhoa_LR * (left ? s.H_l : s.H_r);
    phi = rhoa_LR * magSf;
    phiUp = rhoaU_LR * magSf + p_1_2 * Sf;
    phiEp = rhoah_LR * magSf + p_1_2 * mrf * magSf;
}

inline FaceLR gatherLR(const Recon& r, int f, bool needC)
{
    FaceLR s;
    s.rho_l = r.rho_l[f]; s.rho_r = r.rho_r[f]; s.p_l = r.p_l[f]; s.p_r = r.p_r[f];
    s.U_l = {r.Ux_l[f], r.Uy_l[f], r.Uz_l[f]}; s.U_r = {r.Ux_r[f], r.Uy_r[f], r.Uz_r[f]};
    s.c_l = needC ? r.c_l[f] : 0; s.c_r = needC ? r.c_r[f] : 0;
    s.E_l = r.E_l[f]; s.E_r = r.E_r[f]; s.H_l = r.H_l[f]; s.H_r = r.H_r[f];
    return s;
}

}  // namespace

// convectiveFluxScheme::calcFlux
// ---------------------------------------------------------------------------------------------- Rusanov
// No reference counterpart (src/Make/files:47-51 lists HLLC, ROE, AUSM+up only): the local Lax-Friedrichs flux of the limited
// states in the relative frame — the flux whose frozen-lambda linearisation is the reference's approximate Jacobian
// (convectiveFluxScheme.C:402-546).  Pinned by known answers only (tests/test_oracle_kat.py).
void fluxRusanov(const Ctx&, const FaceLR& s, V3 Sf, double magSf, double mrf, double& phi, V3& phiUp, double& phiEp)
{
    const V3 n = Sf / magSf;
    const double uL = (s.U_l & n) - mrf, uR = (s.U_r & n) - mrf;
    const double lam = std::max(std::fabs(uL) + s.c_l, std::fabs(uR) + s.c_r);
    const double mL = s.rho_l * uL, mR = s.rho_r * uR;
    phi = (0.5 * (mL + mR) - 0.5 * lam * (s.rho_r - s.rho_l)) * magSf;
    phiUp = (0.5 * ((mL * s.U_l + s.p_l * n) + (mR * s.U_r + s.p_r * n)) - (0.5 * lam) * (s.rho_r * s.U_r - s.rho_l * s.U_l)) * magSf;
    phiEp = (0.5 * ((mL * s.H_l + s.p_l * mrf) + (mR * s.H_r + s.p_r * mrf)) - 0.5 * lam * (s.rho_r * s.E_r - s.rho_l * s.E_l)) * magSf;
}

// ---------------------------------------------------------------------------------------------- the face loop
void calcFlux(Ctx& c)
{
    const Mesh& m = c.m;
    const int scheme = c.sch.flux_scheme;
    Derived d;
    derivedFields(c, d, scheme == ICSB200_FLUX_AUSMPLUSUP ? 2 : 0);
    Recon r;
    const bool needC = scheme != ICSB200_FLUX_ROE;
    reconstructAll(c, d, r, needC);
    auto em = emptyMask(m);
    c.phi.assign(m.FT, 0.0); c.phiUp.assign(3 * (size_t)m.FT, 0.0); c.phiEp.assign(m.FT, 0.0);
    for (int f = 0; f < m.FT; f++) {
        if (!faceActive(m, f, em)) continue;
        FaceLR s = gatherLR(r, f, needC);
        V3 Sf = {m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]};
        double phi, phiEp;
        V3 phiUp;
        const double mrf = c.mrfAt(f);  // flux.MRFFaceVelocity() (outerLoop.H:18-21)
        if (scheme == ICSB200_FLUX_HLLC) fluxHLLC(c, s, Sf, m.magSf[f], mrf, phi, phiUp, phiEp);
        else if (scheme == ICSB200_FLUX_ROE) fluxROE(c, s, Sf, m.magSf[f], mrf, phi, phiUp, phiEp);
        else if (scheme == ICSB200_FLUX_RUSANOV) fluxRusanov(c, s, Sf, m.magSf[f], mrf, phi, phiUp, phiEp);
        else fluxAUSM(c, s, Sf, m.magSf[f], mrf, phi, phiUp, phiEp);
        c.phi[f] = phi; c.phiEp[f] = phiEp;
        c.phiUp[3 * (size_t)f] = phiUp.x; c.phiUp[3 * (size_t)f + 1] = phiUp.y; c.phiUp[3 * (size_t)f + 2] = phiUp.z;
    }
    c.phiValid = true;
}

// fvm::ddt(vf) with 'dualTime rPseudoDeltaT <inner>' (dualTimeDdtScheme.C:111-126): scalar diag, source with nc comps
static void fvmDdt(const Ctx& c, const vecd& vf, const vecd& vf0, const vecd& vf00, int nc, vecd& diag, vecd& source)
{
    const Mesh& m = c.m;
    diag.assign(m.N, 0.0);
    source.assign((size_t)nc * m.N, 0.0);
    if (c.sch.ddt_scheme == ICSB200_DDT_EULER) {
        double rDeltaT = 1.0 / c.sch.delta_t;
        for (int i = 0; i < m.N; i++) {
            diag[i] = rDeltaT * m.V[i];
            for (int d = 0; d < nc; d++) source[(size_t)nc * i + d] = rDeltaT * vf0[(size_t)nc * i + d] * m.V[i];
        }
    } else if (c.sch.ddt_scheme == ICSB200_DDT_BACKWARD) {
        double deltaT = c.sch.delta_t, rDeltaT = 1.0 / deltaT;
        double deltaT0 = (c.timeIndex < 2) ? GREAT : deltaT;  // backwardDdtScheme::deltaT0_(vf)
        double coefft = 1 + deltaT / (deltaT + deltaT0);
        double coefft00 = deltaT * deltaT / (deltaT0 * (deltaT + deltaT0));
        double coefft0 = coefft + coefft00;
        for (int i = 0; i < m.N; i++) {
            diag[i] = (coefft * rDeltaT) * m.V[i];
            for (int d = 0; d < nc; d++)
                source[(size_t)nc * i + d] = rDeltaT * m.V[i] * (coefft0 * vf0[(size_t)nc * i + d] - coefft00 * vf00[(size_t)nc * i + d]);
        }
    }
    for (int i = 0; i < m.N; i++) {
        diag[i] += c.rPseudoDeltaT[i] * m.V[i];
        for (int d = 0; d < nc; d++) source[(size_t)nc * i + d] += c.rPseudoDeltaT[i] * vf[(size_t)nc * i + d] * m.V[i];
    }
}


// ---------------------------------------------------------------------------------------------- viscous residual
// residualsUpdate.H:16-43 (laminar: muEff = mu, alphaEff = gamma mu / Pr constant fields):
//   tauMC = muEff*dev2(T(fvc::grad(U)));  rhoUR += fvc::laplacian(muEff,U) + fvc::div(tauMC);
//   sigmaDotU = (interpolate(muEff)*interpolate(grad(U)) + interpolate(tauMC)) & interpolate(U);
//   rhoER += fvc::div(sigmaDotU & Sf) + fvc::laplacian(alphaEff, eCalc),  eCalc = rhoE/rho - 0.5 magSqr(U)
// OpenFOAM semantics restated (fvSchemes: laplacianSchemes `Gauss linear corrected`, gradSchemes / divSchemes
// `Gauss linear`, snGradSchemes `corrected`):
//   fvc::laplacian(g, vf) = fvc::div(interpolate(g) * snGrad(vf) * magSf)
//   snGrad(vf)_f = nonOrthDeltaCoeffs (vf_N - vf_P) + nonOrthCorrectionVectors & interpolate(grad(vf.component))
//   nonOrthCorrectionVectors = n - delta * nonOrthDeltaCoeffs (zero on non-coupled patches)
//   patch snGrad: fvPatchField::snGrad() of the patch field type (below); eCalc has calculated patches
//   grad(U) patch values: gaussGrad::correctBoundaryConditions: gb = gP + n (snGrad(U)_b - n & gP)
namespace {

// fvPatchField<vector>::snGrad() of the U patch field on face f (deltaCoeffs = patch.deltaCoeffs() = 1/|delta|)
void snGradUPatch(const Ctx& c, int pi, int f, double out[3])
{
    const Mesh& m = c.m;
    const int b = f - m.F, o = m.owner[f], s = m.N + b;
    const double dc = m.deltaCoeffs[f];
    const double* Ui = &c.U[3 * (size_t)o];
    const double* Ub = &c.U[3 * (size_t)s];
    switch (c.bc[pi][ICSB200_FIELD_U].kind) {
        case ICSB200_BC_FIXEDVALUE:
        case ICSB200_BC_PRESSUREINLETOUTLETVELOCITY:  // directionMixed::snGrad = (normalValue + transformGradValue - pif)*deltaCoeffs
            for (int d = 0; d < 3; d++) out[d] = dc * (Ub[d] - Ui[d]);
            break;
        case ICSB200_BC_SLIP: {  // basicSymmetryFvPatchField::snGrad = (transform(I - 2.0*sqr(nHat), pif) - pif)*(deltaCoeffs/2.0)
            double n[3];
            for (int d = 0; d < 3; d++) n[d] = m.Sf[3 * f + d] / m.magSf[f];
            double xx = 1.0 - 2.0 * (n[0] * n[0]), xy = 0.0 - 2.0 * (n[0] * n[1]), xz = 0.0 - 2.0 * (n[0] * n[2]);
            double yy = 1.0 - 2.0 * (n[1] * n[1]), yz = 0.0 - 2.0 * (n[1] * n[2]), zz = 1.0 - 2.0 * (n[2] * n[2]);
            double t[3] = {xx * Ui[0] + xy * Ui[1] + xz * Ui[2], xy * Ui[0] + yy * Ui[1] + yz * Ui[2], xz * Ui[0] + yz * Ui[1] + zz * Ui[2]};
            for (int d = 0; d < 3; d++) out[d] = (t[d] - Ui[d]) * (dc / 2.0);
            break;
        }
        case ICSB200_BC_INLETOUTLET: {  // mixedFvPatchField::snGrad, refGrad = 0; valueFraction as frozen at the last evaluation
            const double vfrac = 1.0 - c.vicU[3 * (size_t)b];
            const BC& bc = c.bc[pi][ICSB200_FIELD_U];
            for (int d = 0; d < 3; d++) out[d] = vfrac * (bc.P(f - c.m.patches[pi].start)[d] - Ui[d]) * dc + (1.0 - vfrac) * 0.0;
            break;
        }
        default:  // zeroGradient
            out[0] = out[1] = out[2] = 0.0;
    }
}

inline void dev2T(const double g[9] /* grad(U): g[3*i+j] = d_i U_j */, double mu, double tau[9])
{
    // mu * dev2(T(g)),  dev2(A) = A - (2/3) tr(A) I
    double A[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) A[3 * i + j] = g[3 * j + i];
    const double tr = A[0] + A[4] + A[8];
    const double sph = (2.0 / 3.0) * tr;
    for (int k = 0; k < 9; k++) tau[k] = A[k];
    tau[0] = A[0] - sph; tau[4] = A[4] - sph; tau[8] = A[8] - sph;
    for (int k = 0; k < 9; k++) tau[k] = mu * tau[k];
}

}  // namespace

void viscousResidual(Ctx& c, vecd& rhoUR, vecd& rhoER)
{
    const Mesh& m = c.m;
    const size_t n = (size_t)m.N + m.NB;
    // muEff / alphaEff: laminar constants, or the turbulence model's fields (orc_transport_set)
    vecd muV(n, c.mu), alV(n, c.gamma * (c.mu / c.Pr));
    if (!c.muEffField.empty()) { muV = c.muEffField; alV = c.alphaEffField; syncCoupled(c, muV, 1); syncCoupled(c, alV, 1); }
    // gradients of the components of U and of eCalc (cells + coupled boundary slots)
    vecd comp(n), gU[3], eCalc(n), gE;
    for (int d = 0; d < 3; d++) {
        for (size_t i = 0; i < n; i++) comp[i] = c.U[3 * i + d];
        gradGauss(c, comp, gU[d]);
    }
    for (size_t i = 0; i < n; i++) {
        const double* u = &c.U[3 * i];
        eCalc[i] = c.rho[i] != 0.0 ? c.rhoE[i] / c.rho[i] - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) : 0.0;
    }
    syncCoupled(c, eCalc, 1);  // coupled patches: patchNeighbourField of eCalc's internal field (AMI: interpolated cell values)
    gradGauss(c, eCalc, gE);
    auto gradUOf = [&](size_t i, double g[9]) {  // g[3*i+j] = d_i U_j
        for (int d = 0; d < 3; d++) for (int j = 0; j < 3; j++) g[3 * d + j] = gU[j][3 * i + d];
    };
    const size_t FT = m.FT;
    vecd lapU(3 * FT, 0.0), divTau(3 * FT, 0.0), sig(FT, 0.0), lapE(FT, 0.0);
    auto em = emptyMask(m);
    // rot != nullptr: face of a rotational cyclic patch; nbCell = the neighbour patch's face cell.  The boundary slot Ns then
    // holds transform(forwardT, grad(U_j)) per component, i.e. (T & gradU); cyclicFvPatchField<tensor>::patchNeighbourField
    // of grad(U) and of tauMC is transform(forwardT, tensor) = (T & t) & T.T() (originalOFFiles/.../cyclicFvPatchField.C:130-190)
    // ami != nullptr: face i of a cyclicAMI patch — tauMC is a cell field, so its patchNeighbourField is the AMI interpolation
    // of the neighbour cells' tensors (result = 0; result += w_k tau_k), not the tensor of the interpolated gradient
    auto faceCoupledOrInternal = [&](int f, int P, size_t Ns, const double* dvec, bool coupled, const double* rot, int nbCell, const Patch* ami) {
        const double w = m.w[f], magSf = m.magSf[f], dcn = m.nonOrthDeltaCoeffs[f];
        double nf[3], corr[3];
        for (int d = 0; d < 3; d++) nf[d] = m.Sf[3 * f + d] / magSf;
        for (int d = 0; d < 3; d++) corr[d] = nf[d] - dvec[d] * dcn;
        auto lin = [&](double a, double b) { return coupled ? w * a + (1.0 - w) * b : w * (a - b) + b; };
        double gP[9], gN[9], gf[9], tP[9], tN[9], tf[9];
        gradUOf(P, gP); gradUOf(Ns, gN);
        dev2T(gP, muV[P], tP);
        if (rot || ami) {
            double h[9], tr[9], hr[9];
            if (rot) {
                for (int i = 0; i < 3; i++)           // (T & gradU) & T.T(): the slot already holds the left product
                    for (int j = 0; j < 3; j++) h[3 * i + j] = gN[3 * i] * rot[3 * j] + gN[3 * i + 1] * rot[3 * j + 1] + gN[3 * i + 2] * rot[3 * j + 2];
                for (int k = 0; k < 9; k++) gN[k] = h[k];
            }
            double gRaw[9];
            if (ami) {
                const Patch& q = m.patches[ami->nbrPatch];
                const int i = f - ami->start;
                for (int k = 0; k < 9; k++) tr[k] = 0.0;
                for (int a = ami->amiStart[i]; a < ami->amiStart[i + 1]; a++) {
                    const int cellK = m.owner[q.start + ami->amiFace[a]];
                    double tk[9];
                    gradUOf((size_t)cellK, gRaw);
                    dev2T(gRaw, muV[cellK], tk);
                    for (int k = 0; k < 9; k++) tr[k] += ami->amiWeight[a] * tk[k];
                }
            } else {
                gradUOf((size_t)nbCell, gRaw);         // tauMC is a cell field: rotate the neighbour CELL's tensor
                dev2T(gRaw, muV[Ns], tr);
            }
            if (!rot) { for (int k = 0; k < 9; k++) tN[k] = tr[k]; }
            else {
            for (int i = 0; i < 3; i++)
                for (int l = 0; l < 3; l++) hr[3 * i + l] = rot[3 * i] * tr[l] + rot[3 * i + 1] * tr[3 + l] + rot[3 * i + 2] * tr[6 + l];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) tN[3 * i + j] = hr[3 * i] * rot[3 * j] + hr[3 * i + 1] * rot[3 * j + 1] + hr[3 * i + 2] * rot[3 * j + 2];
            }
        } else
            dev2T(gN, muV[Ns], tN);
        for (int k = 0; k < 9; k++) { gf[k] = lin(gP[k], gN[k]); tf[k] = lin(tP[k], tN[k]); }
        const double muf = lin(muV[P], muV[Ns]), alf = lin(alV[P], alV[Ns]);
        double Uf[3];
        for (int j = 0; j < 3; j++) Uf[j] = lin(c.U[3 * (size_t)P + j], c.U[3 * Ns + j]);
        for (int j = 0; j < 3; j++) {
            // component j: gradient of U_j at the face = column j of gf
            const double snG = dcn * (c.U[3 * Ns + j] - c.U[3 * (size_t)P + j]) + (corr[0] * gf[j] + corr[1] * gf[3 + j] + corr[2] * gf[6 + j]);
            lapU[3 * (size_t)f + j] = muf * snG * magSf;
            divTau[3 * (size_t)f + j] = m.Sf[3 * f] * tf[j] + m.Sf[3 * f + 1] * tf[3 + j] + m.Sf[3 * f + 2] * tf[6 + j];
        }
        double sd[3];
        for (int i = 0; i < 3; i++) {
            const double a0 = muf * gf[3 * i] + tf[3 * i], a1 = muf * gf[3 * i + 1] + tf[3 * i + 1], a2 = muf * gf[3 * i + 2] + tf[3 * i + 2];
            sd[i] = a0 * Uf[0] + a1 * Uf[1] + a2 * Uf[2];
        }
        sig[f] = sd[0] * m.Sf[3 * f] + sd[1] * m.Sf[3 * f + 1] + sd[2] * m.Sf[3 * f + 2];
        double gEf[3];
        for (int d = 0; d < 3; d++) gEf[d] = lin(gE[3 * (size_t)P + d], gE[3 * Ns + d]);
        const double snE = dcn * (eCalc[Ns] - eCalc[P]) + (corr[0] * gEf[0] + corr[1] * gEf[1] + corr[2] * gEf[2]);
        lapE[f] = alf * snE * magSf;
    };
    for (int f = 0; f < m.F; f++) {
        const int P = m.owner[f], N = m.neighbour[f];
        const double dvec[3] = {m.C[3 * (size_t)N] - m.C[3 * (size_t)P], m.C[3 * (size_t)N + 1] - m.C[3 * (size_t)P + 1], m.C[3 * (size_t)N + 2] - m.C[3 * (size_t)P + 2]};
        faceCoupledOrInternal(f, P, (size_t)N, dvec, false, nullptr, -1, nullptr);
    }
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        const Patch& p = m.patches[pi];
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            if (!faceActive(m, f, em)) continue;
            const int b = f - m.F, P = m.owner[f];
            const size_t s = (size_t)m.N + b;
            if (m.coupled(p)) {
                const bool rot = p.rotational && (p.kind == ICSB200_CYCLIC || p.kind == ICSB200_CYCLICAMI);
                const bool isAmi = p.kind == ICSB200_CYCLICAMI;
                faceCoupledOrInternal(f, P, s, &m.dCoupled[3 * (size_t)b], true, rot ? p.forwardT : nullptr,
                                      (rot && !isAmi) ? m.owner[m.patches[p.nbrPatch].start + (f - p.start)] : -1, isAmi ? &p : nullptr);
                continue;
            }
            const double magSf = m.magSf[f];
            double nf[3], sn[3], gP[9], gb[9], tb[9];
            for (int d = 0; d < 3; d++) nf[d] = m.Sf[3 * f + d] / magSf;
            snGradUPatch(c, (int)pi, f, sn);
            gradUOf(P, gP);
            // gaussGrad::correctBoundaryConditions: gb = gP + n * (snGrad - (n & gP))
            for (int j = 0; j < 3; j++) {
                const double ng = nf[0] * gP[j] + nf[1] * gP[3 + j] + nf[2] * gP[6 + j];
                for (int i = 0; i < 3; i++) gb[3 * i + j] = gP[3 * i + j] + nf[i] * (sn[j] - ng);
            }
            const double mu = muV[s], alphaEff = alV[s];  // patch values of muEff / alphaEff
            dev2T(gb, mu, tb);
            const double* Ub = &c.U[3 * s];
            for (int j = 0; j < 3; j++) {
                lapU[3 * (size_t)f + j] = mu * sn[j] * magSf;
                divTau[3 * (size_t)f + j] = m.Sf[3 * f] * tb[j] + m.Sf[3 * f + 1] * tb[3 + j] + m.Sf[3 * f + 2] * tb[6 + j];
            }
            double sd[3];
            for (int i = 0; i < 3; i++) {
                const double a0 = mu * gb[3 * i] + tb[3 * i], a1 = mu * gb[3 * i + 1] + tb[3 * i + 1], a2 = mu * gb[3 * i + 2] + tb[3 * i + 2];
                sd[i] = a0 * Ub[0] + a1 * Ub[1] + a2 * Ub[2];
            }
            sig[f] = sd[0] * m.Sf[3 * f] + sd[1] * m.Sf[3 * f + 1] + sd[2] * m.Sf[3 * f + 2];
            lapE[f] = alphaEff * (m.deltaCoeffs[f] * (eCalc[s] - eCalc[P])) * magSf;
        }
    }
    vecd dv;
    surfaceIntegrate(c, lapU, 3, dv);
    for (size_t i = 0; i < dv.size(); i++) rhoUR[i] += dv[i];
    surfaceIntegrate(c, divTau, 3, dv);
    for (size_t i = 0; i < dv.size(); i++) rhoUR[i] += dv[i];
    surfaceIntegrate(c, sig, 1, dv);
    for (size_t i = 0; i < dv.size(); i++) rhoER[i] += dv[i];
    surfaceIntegrate(c, lapE, 1, dv);
    for (size_t i = 0; i < dv.size(); i++) rhoER[i] += dv[i];
}

// residualsUpdate.H:1-83
void residualsUpdate(Ctx& c)
{
    const Mesh& m = c.m;
    vecd rhoR, rhoUR, rhoER;
    surfaceIntegrate(c, c.phi, 1, rhoR);
    surfaceIntegrate(c, c.phiUp, 3, rhoUR);
    surfaceIntegrate(c, c.phiEp, 1, rhoER);
    for (auto& v : rhoR) v = -v;
    for (auto& v : rhoUR) v = -v;
    for (auto& v : rhoER) v = -v;
    if (c.mu > 0) viscousResidual(c, rhoUR, rhoER);  // if (!inviscid)  (createFields.H:37-45)
    if (c.sch.ddt_scheme != ICSB200_DDT_STEADY) {
        vecd dg, sr;
        fvmDdt(c, c.rho, c.rho0, c.rho00, 1, dg, sr);
        for (int i = 0; i < m.N; i++) rhoR[i] -= (dg[i] * c.rho[i] - sr[i]) / m.V[i];
        fvmDdt(c, c.rhoU, c.rhoU0, c.rhoU00, 3, dg, sr);
        for (int i = 0; i < m.N; i++)
            for (int d = 0; d < 3; d++) rhoUR[3 * (size_t)i + d] -= (dg[i] * c.rhoU[3 * (size_t)i + d] - sr[3 * (size_t)i + d]) / m.V[i];
        fvmDdt(c, c.rhoE, c.rhoE0, c.rhoE00, 1, dg, sr);
        for (int i = 0; i < m.N; i++) rhoER[i] -= (dg[i] * c.rhoE[i] - sr[i]) / m.V[i];
    }
    c.srcRho.resize(m.N); c.srcRhoU.resize(3 * (size_t)m.N); c.srcRhoE.resize(m.N);
    for (int i = 0; i < m.N; i++) {
        c.srcRho[i] = rhoR[i] * m.V[i];
        for (int d = 0; d < 3; d++) c.srcRhoU[3 * (size_t)i + d] = rhoUR[3 * (size_t)i + d] * m.V[i];
        c.srcRhoE[i] = rhoER[i] * m.V[i];
    }
    c.srcMrfApplied = false;
}

// lambda = interpolate(sqrt(gamma/psi)) + mag((interpolate(U) & Sf/magSf) - MRFFaceVelocity)
static void spectralRadius(Ctx& c, vecd& lambda)
{
    const Mesh& m = c.m;
    size_t n = (size_t)m.N + m.NB;
    vecd cc(n), comp(n), cf, uf[3];
    for (size_t i = 0; i < n; i++) cc[i] = std::sqrt(c.gamma / c.psi[i]);
    for (auto& p : m.patches) if (m.empty(p)) for (int f = p.start; f < p.start + p.size; f++) cc[m.N + f - m.F] = 0;
    syncCoupled(c, cc, 1);  // patchNeighbourField of sqrt(gamma/psi) (see derivedFields)
    applyPhaseLag(c, cc, 1, LAG_C1);
    interpolateLinear(c, cc, cf);
    for (int d = 0; d < 3; d++) {
        for (size_t i = 0; i < n; i++) comp[i] = c.U[3 * i + d];
        interpolateLinear(c, comp, uf[d]);
    }
    lambda.assign(m.FT, 0.0);
    auto em = emptyMask(m);
    for (int f = 0; f < m.FT; f++) {
        if (!faceActive(m, f, em)) continue;
        V3 nn = V3{m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]} / m.magSf[f];
        V3 u = {uf[0][f], uf[1][f], uf[2][f]};
        lambda[f] = cf[f] + std::fabs((u & nn) - c.mrfAt(f));
    }
}

// setCoAndDeltaT.H:3-37 — switched evolution relaxation
void serUpdate(Ctx& c)
{
    if (c.haveInitRes) {
        if (!c.firstIter && c.havePrevRes) {
            const icsb200_residuals &ir = c.initRes, &pr = c.prevRes;
            double normInit = std::sqrt(sqr(ir.s_init[0]) + sqr(ir.s_init[1]) + (ir.v_init[0] * ir.v_init[0] + ir.v_init[1] * ir.v_init[1] + ir.v_init[2] * ir.v_init[2]));
            double normPrev = std::sqrt(sqr(pr.s_init[0]) + sqr(pr.s_init[1]) + (pr.v_init[0] * pr.v_init[0] + pr.v_init[1] * pr.v_init[1] + pr.v_init[2] * pr.v_init[2]));
            double coNumRatio = normPrev / normInit;
            coNumRatio = std::max(std::min(coNumRatio, c.sch.pseudo_co_num_max_incr), c.sch.pseudo_co_num_min_decr);
            if (c.sch.local_timestepping) {
                for (auto& v : c.pseudoCoField) { v *= coNumRatio; v = std::max(std::min(v, c.sch.pseudo_co_num_max), c.sch.pseudo_co_num_min); }
            } else {
                c.pseudoCoNum *= coNumRatio;
                c.pseudoCoNum = std::max(std::min(c.pseudoCoNum, c.sch.pseudo_co_num_max), c.sch.pseudo_co_num_min);
            }
        }
        c.prevRes = c.initRes;
        c.havePrevRes = true;
    }
}

// setCoAndDeltaT.H:39-173 — pseudo time step from the current pseudo Courant number
void pseudoDeltaT(Ctx& c)
{
    const Mesh& m = c.m;
    vecd lambda;
    spectralRadius(c, lambda);
    if (c.sch.local_timestepping) {
        std::fill(c.rPseudoDeltaT.begin(), c.rPseudoDeltaT.end(), 0.0);
        for (int f = 0; f < m.F; f++) {
            double frdt = m.nonOrthDeltaCoeffs[f] * lambda[f];
            int o = m.owner[f], n = m.neighbour[f];
            c.rPseudoDeltaT[o] = std::max(c.rPseudoDeltaT[o], frdt);
            c.rPseudoDeltaT[n] = std::max(c.rPseudoDeltaT[n], frdt);
        }
        for (auto& p : m.patches) {
            if (m.coupled(p)) {
                for (int f = p.start; f < p.start + p.size; f++) {
                    int o = m.owner[f];
                    c.rPseudoDeltaT[o] = std::max(c.rPseudoDeltaT[o], m.nonOrthDeltaCoeffs[f] * lambda[f]);
                }
            } else if (p.kind == ICSB200_WALL) {
                for (int f = p.start; f < p.start + p.size; f++) {
                    int o = m.owner[f];
                    V3 nn = V3{m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]} / m.magSf[f];
                    V3 u = {c.U[3 * (size_t)o], c.U[3 * (size_t)o + 1], c.U[3 * (size_t)o + 2]};
                    double pLambda = 0.5 * m.nonOrthDeltaCoeffs[f] * (std::sqrt(c.gamma / c.psi[o]) + std::fabs((u & nn) - c.mrfAt(f)));
                    c.rPseudoDeltaT[o] = std::max(c.rPseudoDeltaT[o], pLambda);
                }
            }
        }
        for (int i = 0; i < m.N; i++) c.rPseudoDeltaT[i] /= c.pseudoCoField[i];
    } else {
        double mx = -VGREAT;
        auto em = emptyMask(m);
        for (int f = 0; f < m.FT; f++) if (faceActive(m, f, em)) mx = std::max(mx, m.deltaCoeffs[f] * lambda[f]);
        mx = c.comm->max(mx);
        for (int i = 0; i < m.N; i++) c.rPseudoDeltaT[i] = mx / c.pseudoCoNum;
    }
}

void setCoAndDeltaT(Ctx& c)
{
    serUpdate(c);
    pseudoDeltaT(c);
}

// outerLoop.H:61-64
void computeDdtCoeff(Ctx& c)
{
    const Mesh& m = c.m;
    vecd d1, d2, d3, s;
    fvmDdt(c, c.rho, c.rho0, c.rho00, 1, d1, s);
    fvmDdt(c, c.rhoU, c.rhoU0, c.rhoU00, 3, d2, s);
    fvmDdt(c, c.rhoE, c.rhoE0, c.rhoE00, 1, d3, s);
    c.ddtCoeff.resize(m.N);
    for (int i = 0; i < m.N; i++) c.ddtCoeff[i] = std::max(std::max(d1[i], d2[i]), d3[i]) / m.V[i];
}

// ---------------------------------------------------------------------------------------------- Jacobian
namespace {

void blkInit(Blk& b, int nc, const Mesh& m)
{
    b.nc = nc; b.exists = true; b.hasOff = false; b.hasInt = false;
    b.diag.assign((size_t)nc * m.N, 0.0);
    b.upper.clear(); b.lower.clear(); b.intUpper.clear(); b.intLower.clear();
    b.source.assign((size_t)(nc == 9 ? 3 : (nc == 3 ? 0 : 1)) * m.N, 0.0);
}

void ensureOff(Blk& b, const Mesh& m)
{
    if (!b.hasOff) { b.upper.assign((size_t)b.nc * m.F, 0.0); b.lower.assign((size_t)b.nc * m.F, 0.0); b.hasOff = true; }
    if (!b.hasInt) { b.intUpper.assign((size_t)b.nc * m.NB, 0.0); b.intLower.assign((size_t)b.nc * m.NB, 0.0); b.hasInt = true; }
}

// blockFvMatrix::insertBlock (blockFvMatrix.C:211-268): left/right are [nc*FT] face fields
void insertBlock(const Mesh& m, Blk& A, const vecd& left, const vecd& right)
{
    const int nc = A.nc;
    vecd upp((size_t)nc * m.F), low((size_t)nc * m.F), diag((size_t)nc * m.N, 0.0);
    for (int f = 0; f < m.F; f++)
        for (int k = 0; k < nc; k++) {
            upp[(size_t)nc * f + k] = 0.5 * m.magSf[f] * right[(size_t)nc * f + k];
            low[(size_t)nc * f + k] = -0.5 * m.magSf[f] * left[(size_t)nc * f + k];
        }
    for (int f = 0; f < m.F; f++)  // negSumDiag
        for (int k = 0; k < nc; k++) {
            diag[(size_t)nc * m.owner[f] + k] -= low[(size_t)nc * f + k];
            diag[(size_t)nc * m.neighbour[f] + k] -= upp[(size_t)nc * f + k];
        }
    vecd iu((size_t)nc * m.NB, 0.0), il((size_t)nc * m.NB, 0.0);
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++)
            for (int k = 0; k < nc; k++) {
                iu[(size_t)nc * (f - m.F) + k] = 0.5 * m.magSf[f] * right[(size_t)nc * f + k];
                il[(size_t)nc * (f - m.F) + k] = -0.5 * m.magSf[f] * left[(size_t)nc * f + k];
            }
        if (m.coupled(p))
            for (int f = p.start; f < p.start + p.size; f++)
                for (int k = 0; k < nc; k++) diag[(size_t)nc * m.owner[f] + k] -= il[(size_t)nc * (f - m.F) + k];
    }
    ensureOff(A, m);
    for (size_t i = 0; i < diag.size(); i++) A.diag[i] += diag[i];
    for (size_t i = 0; i < upp.size(); i++) { A.upper[i] += upp[i]; A.lower[i] += low[i]; }
    for (size_t i = 0; i < iu.size(); i++) { A.intUpper[i] += iu[i]; A.intLower[i] += il[i]; }
}

// blockFvMatrix::insertDissipationBlock (blockFvMatrix.C:271-326) for a scalar face field times identity pattern
// 'pattern' lists the coefficient slots that receive lambda (scalar: {0}; tensor: {0,4,8})
void insertDissipationBlock(const Mesh& m, Blk& A, const vecd& diss, std::initializer_list<int> pattern)
{
    const int nc = A.nc;
    vecd upp((size_t)nc * m.F, 0.0), low((size_t)nc * m.F, 0.0), diag((size_t)nc * m.N, 0.0);
    for (int f = 0; f < m.F; f++)
        for (int k : pattern) {
            upp[(size_t)nc * f + k] = 0.5 * m.magSf[f] * diss[f];
            low[(size_t)nc * f + k] = 0.5 * m.magSf[f] * diss[f];
        }
    for (int f = 0; f < m.F; f++)
        for (int k : pattern) {
            diag[(size_t)nc * m.owner[f] + k] -= low[(size_t)nc * f + k];
            diag[(size_t)nc * m.neighbour[f] + k] -= upp[(size_t)nc * f + k];
        }
    vecd iu((size_t)nc * m.NB, 0.0), il((size_t)nc * m.NB, 0.0);
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++)
            for (int k : pattern) {
                iu[(size_t)nc * (f - m.F) + k] = 0.5 * m.magSf[f] * diss[f];
                il[(size_t)nc * (f - m.F) + k] = 0.5 * m.magSf[f] * diss[f];
                diag[(size_t)nc * m.owner[f] + k] -= il[(size_t)nc * (f - m.F) + k];  // physical boundaries included
            }
    }
    ensureOff(A, m);
    for (size_t i = 0; i < diag.size(); i++) A.diag[i] -= diag[i];
    for (size_t i = 0; i < upp.size(); i++) { A.upper[i] -= upp[i]; A.lower[i] -= low[i]; }
    for (size_t i = 0; i < iu.size(); i++) { A.intUpper[i] -= iu[i]; A.intLower[i] -= il[i]; }
}

// fvj::laplacian(sf, geometricOneField) (blockFvOperatorsTemplates.C:499-557), then A -= stab [* I]
void subtractLaplacianOne(const Mesh& m, Blk& A, const vecd& sf, std::initializer_list<int> pattern)
{
    const int nc = A.nc;
    vecd sf2(m.FT, 0.0);
    for (int f = 0; f < m.FT; f++) sf2[f] = sf[f] * m.magSf[f] * m.deltaCoeffs[f];
    vecd diag(m.N, 0.0);
    for (int f = 0; f < m.F; f++) { diag[m.owner[f]] -= sf2[f]; diag[m.neighbour[f]] -= sf2[f]; }
    for (auto& p : m.patches)
        if (m.coupled(p)) for (int f = p.start; f < p.start + p.size; f++) diag[m.owner[f]] -= sf2[f];
    ensureOff(A, m);
    for (int k : pattern) {
        for (int i = 0; i < m.N; i++) A.diag[(size_t)nc * i + k] -= diag[i] * 1.0;
        for (int f = 0; f < m.F; f++) { A.upper[(size_t)nc * f + k] -= sf2[f] * 1.0; A.lower[(size_t)nc * f + k] -= sf2[f] * 1.0; }
        for (auto& p : m.patches) {
            if (m.empty(p)) continue;
            for (int f = p.start; f < p.start + p.size; f++) {
                A.intUpper[(size_t)nc * (f - m.F) + k] -= sf2[f] * 1.0;
                A.intLower[(size_t)nc * (f - m.F) + k] -= sf2[f] * 1.0;
            }
        }
    }
}

// fvj::div(w, sf) for a face scalar already dotted with Sf (blockFvOperatorsTemplates.C:146-201), then A -= mx [* I]
void subtractDivFace(const Mesh& m, Blk& A, const vecd& sf, std::initializer_list<int> pattern)
{
    const int nc = A.nc;
    vecd upp(m.F), low(m.F), diag(m.N, 0.0), iu(m.NB, 0.0), il(m.NB, 0.0);
    for (int f = 0; f < m.F; f++) { upp[f] = sf[f] * (1 - m.w[f]); low[f] = -sf[f] * m.w[f]; }
    for (int f = 0; f < m.F; f++) { diag[m.owner[f]] -= low[f]; diag[m.neighbour[f]] -= upp[f]; }
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) { iu[f - m.F] = sf[f] * (1 - m.w[f]); il[f - m.F] = -sf[f] * m.w[f]; }
        if (m.coupled(p)) for (int f = p.start; f < p.start + p.size; f++) diag[m.owner[f]] -= il[f - m.F];
    }
    ensureOff(A, m);
    for (int k : pattern) {
        for (int i = 0; i < m.N; i++) A.diag[(size_t)nc * i + k] -= diag[i] * 1.0;
        for (int f = 0; f < m.F; f++) { A.upper[(size_t)nc * f + k] -= upp[f] * 1.0; A.lower[(size_t)nc * f + k] -= low[f] * 1.0; }
        for (int b = 0; b < m.NB; b++) { A.intUpper[(size_t)nc * b + k] -= iu[b] * 1.0; A.intLower[(size_t)nc * b + k] -= il[b] * 1.0; }
    }
}

// fvj::laplacian(sf, vf) (blockFvOperatorsTemplates.C:387-450 / :575-637), then A -= mx [* I]: vf is a cell field with nv
// components over cells and boundary slots (boundary = the values STORED on the patch, see storedCoupledValues), sf2 = sf |Sf|
// deltaCoeffs; upper = vf[nei] sf2, lower = vf[own] sf2, negSumDiag, coupled patches: diag[own] -= vf[own] sf2_b
void subtractLaplacianField(const Mesh& m, Blk& A, const vecd& sf2, const vecd& vf, int nv, const int* slots)
{
    const int nc = A.nc;
    vecd upp((size_t)nv * m.F), low((size_t)nv * m.F), diag((size_t)nv * m.N, 0.0), iu((size_t)nv * m.NB, 0.0), il((size_t)nv * m.NB, 0.0);
    for (int f = 0; f < m.F; f++)
        for (int k = 0; k < nv; k++) {
            upp[(size_t)nv * f + k] = vf[(size_t)nv * m.neighbour[f] + k] * sf2[f];
            low[(size_t)nv * f + k] = vf[(size_t)nv * m.owner[f] + k] * sf2[f];
        }
    for (int f = 0; f < m.F; f++)
        for (int k = 0; k < nv; k++) {
            diag[(size_t)nv * m.owner[f] + k] -= low[(size_t)nv * f + k];
            diag[(size_t)nv * m.neighbour[f] + k] -= upp[(size_t)nv * f + k];
        }
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++)
            for (int k = 0; k < nv; k++) {
                iu[(size_t)nv * (f - m.F) + k] = vf[(size_t)nv * (m.N + f - m.F) + k] * sf2[f];
                il[(size_t)nv * (f - m.F) + k] = vf[(size_t)nv * m.owner[f] + k] * sf2[f];
            }
        if (m.coupled(p))
            for (int f = p.start; f < p.start + p.size; f++)
                for (int k = 0; k < nv; k++) diag[(size_t)nv * m.owner[f] + k] -= il[(size_t)nv * (f - m.F) + k];
    }
    ensureOff(A, m);
    // slots[k]: coefficient slot of A receiving component k (vector blocks: 0,1,2; scalar into tensor: one call per 0,4,8 with * 1.0)
    for (int k = 0; k < nv; k++) {
        const int s = slots[k];
        for (int i = 0; i < m.N; i++) A.diag[(size_t)nc * i + s] -= diag[(size_t)nv * i + k];
        for (int f = 0; f < m.F; f++) { A.upper[(size_t)nc * f + s] -= upp[(size_t)nv * f + k]; A.lower[(size_t)nc * f + s] -= low[(size_t)nv * f + k]; }
        for (int b = 0; b < m.NB; b++) { A.intUpper[(size_t)nc * b + s] -= iu[(size_t)nv * b + k]; A.intLower[(size_t)nc * b + s] -= il[(size_t)nv * b + k]; }
    }
}

}  // namespace

// Values STORED on the boundary patches of p, U, T, rho, E as the solver leaves them (updateFields.H:80-104): physical
// patches carry their boundary condition value; a coupledFvPatchField (cyclic, cyclicAMI) stores
// w*internal + (1-w)*patchNeighbourField after evaluate(), a processor patch stores the neighbour values; then T = THE(he(T)),
// psi = 1/(R T), rho = psi p and E = he(p,T) + 0.5 |U|^2 are formed from those.  vf.boundaryField() of the expression fields
// handed to fvj::laplacian reads exactly these.
static void storedBoundaryValues(const Ctx& c, vecd& rho, vecd& U, vecd& E)
{
    const Mesh& m = c.m;
    const size_t n = (size_t)m.N + m.NB;
    rho = c.rho; U = c.U; E.assign(n, 0.0);
    for (size_t i = 0; i < n; i++) {
        const double* u = &c.U[3 * i];
        E[i] = c.Cv * c.T[i] + 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    }
    for (auto& p : m.patches) {
        if (!m.coupled(p)) continue;
        const bool interp = p.kind != ICSB200_PROCESSOR;
        for (int f = p.start; f < p.start + p.size; f++) {
            const size_t s = (size_t)m.N + f - m.F;
            const int o = m.owner[f];
            const double w = m.w[f];
            double pb = c.p[s], Tb = c.T[s], Ub[3] = {c.U[3 * s], c.U[3 * s + 1], c.U[3 * s + 2]};
            if (interp) {
                pb = w * c.p[o] + (1.0 - w) * c.p[s];
                Tb = w * c.T[o] + (1.0 - w) * c.T[s];
                for (int d = 0; d < 3; d++) Ub[d] = w * c.U[3 * (size_t)o + d] + (1.0 - w) * c.U[3 * s + d];
            }
            Tb = std::max(Tb, c.sch.T_min);
            const double eb = c.Cv * Tb;
            Tb = eb / c.Cv;                       // thermo.correct(): T = THE(he, p, T0) on patches that do not fix the value
            const double psib = 1.0 / (c.R * Tb);
            rho[s] = psib * pb;
            for (int d = 0; d < 3; d++) U[3 * s + d] = Ub[d];
            E[s] = c.Cv * Tb + 0.5 * (Ub[0] * Ub[0] + Ub[1] * Ub[1] + Ub[2] * Ub[2]);
        }
    }
}

// viscousFluxScheme::addFluxTerms, LaxFriedrichJacobian false (viscousFluxScheme.C:248-261), and addBoundaryTerms (:120-215)
void fullViscousJacobian(Ctx& c)
{
    const Mesh& m = c.m;
    const size_t n = (size_t)m.N + m.NB;
    Blk& dEnergyByRho = c.blk[2]; Blk& dEnergyByRhoE = c.blk[3]; Blk& dEnergyByRhoU = c.blk[5]; Blk& dMomByRho = c.blk[6];
    Blk& dMomByRhoE = c.blk[7]; Blk& dMomByRhoU = c.blk[8];
    vecd muEff(n, c.mu), alphaEff(n, c.gamma * (c.mu / c.Pr)), muf, alf;
    if (!c.muEffField.empty()) { muEff = c.muEffField; alphaEff = c.alphaEffField; syncCoupled(c, muEff, 1); syncCoupled(c, alphaEff, 1); }
    interpolateLinear(c, muEff, muf);
    interpolateLinear(c, alphaEff, alf);
    auto em = emptyMask(m);
    vecd sf2mu(m.FT, 0.0), sf2al(m.FT, 0.0);
    for (int f = 0; f < m.FT; f++)
        if (faceActive(m, f, em)) { sf2mu[f] = muf[f] * m.magSf[f] * m.deltaCoeffs[f]; sf2al[f] = alf[f] * m.magSf[f] * m.deltaCoeffs[f]; }
    vecd rho, U, E;
    storedBoundaryValues(c, rho, U, E);
    vecd mURho(3 * n), invRho(n), eRho(n);
    for (size_t i = 0; i < n; i++) {
        if (rho[i] == 0.0) { mURho[3 * i] = mURho[3 * i + 1] = mURho[3 * i + 2] = invRho[i] = eRho[i] = 0.0; continue; }  // empty-patch slots
        const double* u = &U[3 * i];
        for (int d = 0; d < 3; d++) mURho[3 * i + d] = -u[d] / rho[i];
        invRho[i] = 1.0 / rho[i];
        eRho[i] = -E[i] / rho[i] + (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]) / rho[i];
    }
    static const int v3[3] = {0, 1, 2}, s0[1] = {0}, s4[1] = {4}, s8[1] = {8};
    subtractLaplacianField(m, dMomByRho, sf2mu, mURho, 3, v3);
    subtractLaplacianField(m, dMomByRhoU, sf2mu, invRho, 1, s0);   // fvj::laplacian(muEff, 1/rho) * tensor::I
    subtractLaplacianField(m, dMomByRhoU, sf2mu, invRho, 1, s4);
    subtractLaplacianField(m, dMomByRhoU, sf2mu, invRho, 1, s8);
    subtractLaplacianField(m, dEnergyByRho, sf2al, eRho, 1, s0);
    subtractLaplacianField(m, dEnergyByRhoU, sf2al, mURho, 3, v3);
    subtractLaplacianField(m, dEnergyByRhoE, sf2al, invRho, 1, s0);
    // ---- addBoundaryTerms: physical patches, gradientInternalCoeffs of U and T
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        auto& p = m.patches[pi];
        if (m.coupled(p) || m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            const int b = f - m.F, iIntCell = m.owner[f];
            const double magSfB = m.magSf[f], cvB = c.Cv;
            const double* uGIC = &c.gicU[3 * (size_t)b];
            const double TGIC = c.gicT[b];
            // dMomFluxdUDiag = -muEff_b*uGIC*magSfB ; dEnergyFluxdT = -alphaEff_b*cvB*TGIC*magSfB
            double dMomFluxdU[3];
            for (int d = 0; d < 3; d++) dMomFluxdU[d] = -muf[f] * uGIC[d] * magSfB;
            const double dEnergyFluxdT = -alf[f] * cvB * TGIC * magSfB;
            const double rhoI = c.rho[iIntCell], cvI = c.Cv;
            const V3 UI = {c.U[3 * (size_t)iIntCell], c.U[3 * (size_t)iIntCell + 1], c.U[3 * (size_t)iIntCell + 2]};
            const double EI = c.Cv * c.T[iIntCell] + 0.5 * magSqr(UI);
            const V3 dUdRho = -1.0 * UI / rhoI;
            const double dTdRho = -1.0 / (cvI * rhoI) * (EI - magSqr(UI));
            const double dUdRhoU = 1.0 / rhoI * 1.0;
            const V3 dTdRhoU = -1.0 * UI / (cvI * rhoI);
            const double dTdRhoE = 1.0 / (cvI * rhoI);
            const double du[3] = {dUdRho.x, dUdRho.y, dUdRho.z};
            for (int d = 0; d < 3; d++) {
                dMomByRho.diag[3 * (size_t)iIntCell + d] += dMomFluxdU[d] * du[d];          // tensor (diagonal) & vector
                dMomByRhoU.diag[9 * (size_t)iIntCell + 4 * d] += dMomFluxdU[d] * dUdRhoU;    // tensor & sphericalTensor
                dMomByRhoE.diag[3 * (size_t)iIntCell + d] += dMomFluxdU[d] * 0.0;            // dUdRhoE = 0
            }
            dEnergyByRho.diag[iIntCell] += dEnergyFluxdT * dTdRho;
            dEnergyByRhoU.diag[3 * (size_t)iIntCell] += dEnergyFluxdT * dTdRhoU.x;
            dEnergyByRhoU.diag[3 * (size_t)iIntCell + 1] += dEnergyFluxdT * dTdRhoU.y;
            dEnergyByRhoU.diag[3 * (size_t)iIntCell + 2] += dEnergyFluxdT * dTdRhoU.z;
            dEnergyByRhoE.diag[iIntCell] += dEnergyFluxdT * dTdRhoE;
        }
    }
}

// convectiveFluxScheme::createConvectiveJacobian + viscousFluxScheme::createViscousJacobian (LF branch)
void createJacobian(Ctx& c)
{
    const Mesh& m = c.m;
    // coupledMatrix eqSystem(mesh, 2, 1, true)  (outerLoop.H:53) — fresh every iteration
    Blk& dContByRho = c.blk[0]; Blk& dContByRhoE = c.blk[1]; Blk& dEnergyByRho = c.blk[2]; Blk& dEnergyByRhoE = c.blk[3];
    Blk& dContByRhoU = c.blk[4]; Blk& dEnergyByRhoU = c.blk[5]; Blk& dMomByRho = c.blk[6]; Blk& dMomByRhoE = c.blk[7];
    Blk& dMomByRhoU = c.blk[8];
    blkInit(dContByRho, 1, m); blkInit(dContByRhoE, 1, m); blkInit(dEnergyByRho, 1, m); blkInit(dEnergyByRhoE, 1, m);
    blkInit(dContByRhoU, 3, m); blkInit(dEnergyByRhoU, 3, m); blkInit(dMomByRho, 3, m); blkInit(dMomByRhoE, 3, m);
    blkInit(dMomByRhoU, 9, m);

    // ---- addFluxTerms (convectiveFluxScheme.C:374-484)
    Derived d;
    derivedFields(c, d, 1);
    vecd Ux_l, Ux_r, Uy_l, Uy_r, Uz_l, Uz_r, E_l, E_r;
    interpolateLimitedLR(c, d.Ux, c.sch.limiter_U, Ux_l, Ux_r);
    interpolateLimitedLR(c, d.Uy, c.sch.limiter_U, Uy_l, Uy_r);
    interpolateLimitedLR(c, d.Uz, c.sch.limiter_U, Uz_l, Uz_r);
    interpolateLimitedLR(c, d.E, c.sch.limiter_T, E_l, E_r);
    const size_t FT = m.FT;
    vecd nL(3 * FT, 0), momRho_l(3 * FT, 0), momRho_r(3 * FT, 0), momRhoU_l(9 * FT, 0), momRhoU_r(9 * FT, 0), momRhoE_l(3 * FT, 0),
        momRhoE_r(3 * FT, 0), enRho_l(FT, 0), enRho_r(FT, 0), enRhoU_l(3 * FT, 0), enRhoU_r(3 * FT, 0), enRhoE_l(FT, 0), enRhoE_r(FT, 0);
    auto em = emptyMask(m);
    auto put3 = [](vecd& a, size_t f, V3 v) { a[3 * f] = v.x; a[3 * f + 1] = v.y; a[3 * f + 2] = v.z; };
    for (size_t f = 0; f < FT; f++) {
        if (!faceActive(m, (int)f, em)) continue;
        double gamma_interp = c.gamma;
        V3 U_l = {Ux_l[f], Uy_l[f], Uz_l[f]}, U_r = {Ux_r[f], Uy_r[f], Uz_r[f]};
        double theta_l = 0.5 * (gamma_interp - 1) * magSqr(U_l), theta_r = 0.5 * (gamma_interp - 1) * magSqr(U_r);
        double a1_l = gamma_interp * E_l[f] - theta_l, a1_r = gamma_interp * E_r[f] - theta_r;
        double a2_l = gamma_interp - 1, a2_r = gamma_interp - 1;
        V3 n = V3{m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]} / m.magSf[f];
        double projU_l = U_l & n, projU_r = U_r & n;
        put3(nL, f, n);
        put3(momRho_l, f, n * theta_l - U_l * projU_l);
        put3(momRho_r, f, n * theta_r - U_r * projU_r);
        T9 tl = outer(U_l, n) - outer(a2_l * n, U_l) + projU_l * I9;
        T9 tr = outer(U_r, n) - outer(a2_r * n, U_r) + projU_r * I9;
        for (int k = 0; k < 9; k++) { momRhoU_l[9 * f + k] = tl.v[k]; momRhoU_r[9 * f + k] = tr.v[k]; }
        put3(momRhoE_l, f, n * a2_l);
        put3(momRhoE_r, f, n * a2_r);
        enRho_l[f] = projU_l * (theta_l - a1_l);
        enRho_r[f] = projU_r * (theta_r - a1_r);
        put3(enRhoU_l, f, n * a1_l - a2_l * U_l * projU_l);
        put3(enRhoU_r, f, n * a1_r - a2_r * U_r * projU_r);
        enRhoE_l[f] = gamma_interp * projU_l;
        enRhoE_r[f] = gamma_interp * projU_r;
    }
    insertBlock(m, dContByRhoU, nL, nL);
    insertBlock(m, dMomByRho, momRho_l, momRho_r);
    insertBlock(m, dMomByRhoU, momRhoU_l, momRhoU_r);
    insertBlock(m, dMomByRhoE, momRhoE_l, momRhoE_r);
    insertBlock(m, dEnergyByRho, enRho_l, enRho_r);
    insertBlock(m, dEnergyByRhoU, enRhoU_l, enRhoU_r);
    insertBlock(m, dEnergyByRhoE, enRhoE_l, enRhoE_r);
    // (the moving-mesh fvj::div term is identically zero here: static mesh)
    // MRFdivMeshPhi = fvj::div(w, MRFFaceVelocity*magSf) (convectiveFluxScheme.C:477-481)
    if (!c.mrfFaceVel.empty()) {
        vecd sfm(FT, 0.0);
        for (size_t f = 0; f < FT; f++) if (faceActive(m, (int)f, em)) sfm[f] = c.mrfFaceVel[f] * m.magSf[f];
        subtractDivFace(m, dContByRho, sfm, {0});
        subtractDivFace(m, dMomByRhoU, sfm, {0, 4, 8});
        subtractDivFace(m, dEnergyByRhoE, sfm, {0});
    }

    // ---- addDissipationJacobian (convectiveFluxScheme.C:487-534)
    vecd lambdaConv;
    spectralRadius(c, lambdaConv);
    insertDissipationBlock(m, dContByRho, lambdaConv, {0});
    insertDissipationBlock(m, dMomByRhoU, lambdaConv, {0, 4, 8});
    insertDissipationBlock(m, dEnergyByRhoE, lambdaConv, {0});

    // ---- addBoundaryTerms + boundaryJacobian (convectiveFluxScheme.C:47-120, 219-355)
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        auto& p = m.patches[pi];
        if (m.coupled(p) || m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            int s = m.N + f - m.F, iIntCell = m.owner[f];
            const int bf = f - m.F;
            double pVIC = c.vicP[bf], tVIC = c.vicT[bf];
            V3 uVIC = {c.vicU[3 * (size_t)bf], c.vicU[3 * (size_t)bf + 1], c.vicU[3 * (size_t)bf + 2]};
            V3 SfB = {m.Sf[3 * f], m.Sf[3 * f + 1], m.Sf[3 * f + 2]};
            double rhoB = c.rho[s];
            V3 UB = {c.U[3 * (size_t)s], c.U[3 * (size_t)s + 1], c.U[3 * (size_t)s + 2]};
            double UrelBdotSf = UB & SfB;
            UrelBdotSf -= c.mrfAt(f) * m.magSf[f];
            double pB = c.p[s], TB = c.T[s];
            // rhoEn = rho*(he(p,T) + 0.5 magSqr(U)) evaluated on the boundary
            double rhoEB = rhoB * (c.Cv * TB + 0.5 * magSqr(UB));
            double cvB = c.Cv;
            auto cm = [](V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; };
            double dContFluxdp = rhoB / pB * UrelBdotSf * pVIC;
            V3 dContFluxdU = rhoB * cm(SfB, uVIC);
            double dContFluxdT = -rhoB / TB * UrelBdotSf * tVIC;
            V3 dMomFluxdp = (rhoB / pB * UB * UrelBdotSf + SfB) * pVIC;
            T9 dMomFluxdU = outer(rhoB * UB, cm(SfB, uVIC));
            V3 dMomFluxdUDiag = rhoB * UrelBdotSf * uVIC;
            V3 dMomFluxdT = -rhoB / TB * UB * UrelBdotSf * tVIC;
            double dEnergyFluxdp = (rhoEB / pB * UrelBdotSf + (UB & SfB)) * pVIC;
            V3 dEnergyFluxdU = cm(SfB, uVIC) * (rhoEB + pB) + rhoB * UrelBdotSf * cm(UB, uVIC);
            double dEnergyFluxdT = UrelBdotSf * (rhoB * cvB - rhoEB / TB) * tVIC;
            dMomFluxdU.v[0] = dMomFluxdU.v[0] + dMomFluxdUDiag.x;
            dMomFluxdU.v[4] = dMomFluxdU.v[4] + dMomFluxdUDiag.y;
            dMomFluxdU.v[8] = dMomFluxdU.v[8] + dMomFluxdUDiag.z;
            // internal-cell derivatives of (p, U, T) w.r.t. the conserved variables
            double rhoI = c.rho[iIntCell];
            V3 UI = {c.U[3 * (size_t)iIntCell], c.U[3 * (size_t)iIntCell + 1], c.U[3 * (size_t)iIntCell + 2]};
            double rhoEI = rhoI * (c.Cv * c.T[iIntCell] + 0.5 * magSqr(UI));
            double gammaI = c.gamma, cvI = c.Cv;
            double dPdRho = 0.5 * (gammaI - 1) * magSqr(UI);
            V3 dUdRho = -1.0 * UI / rhoI;
            double dTdRho = -1.0 / (cvI * rhoI) * (rhoEI / rhoI - magSqr(UI));
            V3 dPdRhoU = -(gammaI - 1) * UI;
            double dUdRhoU = 1.0 / rhoI * 1.0;  // sphericalTensor ii component
            V3 dTdRhoU = -1.0 * UI / (cvI * rhoI);
            double dPdRhoE = gammaI - 1;
            V3 dUdRhoE = {0, 0, 0};
            double dTdRhoE = 1.0 / (cvI * rhoI);
            auto add3 = [](vecd& a, int i, V3 v) { a[3 * (size_t)i] += v.x; a[3 * (size_t)i + 1] += v.y; a[3 * (size_t)i + 2] += v.z; };
            // vector & sphericalTensor = vector * ii ; tensor & sphericalTensor = tensor * ii
            dContByRho.diag[iIntCell] += dContFluxdp * dPdRho;
            dContByRho.diag[iIntCell] += dContFluxdU & dUdRho;
            dContByRho.diag[iIntCell] += dContFluxdT * dTdRho;
            add3(dContByRhoU.diag, iIntCell, dContFluxdp * dPdRhoU);
            add3(dContByRhoU.diag, iIntCell, dContFluxdU * dUdRhoU);
            add3(dContByRhoU.diag, iIntCell, dContFluxdT * dTdRhoU);
            dContByRhoE.diag[iIntCell] += dContFluxdp * dPdRhoE;
            dContByRhoE.diag[iIntCell] += dContFluxdU & dUdRhoE;
            dContByRhoE.diag[iIntCell] += dContFluxdT * dTdRhoE;
            add3(dMomByRho.diag, iIntCell, dMomFluxdp * dPdRho);
            add3(dMomByRho.diag, iIntCell, dMomFluxdU & dUdRho);
            add3(dMomByRho.diag, iIntCell, dMomFluxdT * dTdRho);
            T9 t1 = outer(dMomFluxdp, dPdRhoU), t2 = dMomFluxdU * dUdRhoU, t3 = outer(dMomFluxdT, dTdRhoU);
            for (int k = 0; k < 9; k++) dMomByRhoU.diag[9 * (size_t)iIntCell + k] += t1.v[k];
            for (int k = 0; k < 9; k++) dMomByRhoU.diag[9 * (size_t)iIntCell + k] += t2.v[k];
            for (int k = 0; k < 9; k++) dMomByRhoU.diag[9 * (size_t)iIntCell + k] += t3.v[k];
            add3(dMomByRhoE.diag, iIntCell, dMomFluxdp * dPdRhoE);
            add3(dMomByRhoE.diag, iIntCell, dMomFluxdU & dUdRhoE);
            add3(dMomByRhoE.diag, iIntCell, dMomFluxdT * dTdRhoE);
            dEnergyByRho.diag[iIntCell] += dEnergyFluxdp * dPdRho;
            dEnergyByRho.diag[iIntCell] += dEnergyFluxdU & dUdRho;
            dEnergyByRho.diag[iIntCell] += dEnergyFluxdT * dTdRho;
            add3(dEnergyByRhoU.diag, iIntCell, dEnergyFluxdp * dPdRhoU);
            add3(dEnergyByRhoU.diag, iIntCell, dEnergyFluxdU * dUdRhoU);
            add3(dEnergyByRhoU.diag, iIntCell, dEnergyFluxdT * dTdRhoU);
            dEnergyByRhoE.diag[iIntCell] += dEnergyFluxdp * dPdRhoE;
            dEnergyByRhoE.diag[iIntCell] += dEnergyFluxdU & dUdRhoE;
            dEnergyByRhoE.diag[iIntCell] += dEnergyFluxdT * dTdRhoE;
        }
    }

    // ---- addTemporalTerms (convectiveFluxScheme.C:358-369)
    for (int i = 0; i < m.N; i++) {
        double diagCoeff = c.ddtCoeff[i] * m.V[i];
        dContByRho.diag[i] += diagCoeff;
        for (int k = 0; k < 9; k++) dMomByRhoU.diag[9 * (size_t)i + k] += diagCoeff * I9.v[k];
        dEnergyByRhoE.diag[i] += diagCoeff;
    }
    // addMRFSource, diagonal part (convectiveFluxScheme.C:130-137)
    if (!c.mrfOmega.empty())
        for (int i = 0; i < m.N; i++) {
            const double ox = c.mrfOmega[3 * (size_t)i], oy = c.mrfOmega[3 * (size_t)i + 1], oz = c.mrfOmega[3 * (size_t)i + 2];
            double* md = &dMomByRhoU.diag[9 * (size_t)i];
            md[1] -= oz * m.V[i]; md[2] += oy * m.V[i]; md[3] += oz * m.V[i];
            md[5] -= ox * m.V[i]; md[6] -= oy * m.V[i]; md[7] += ox * m.V[i];
        }

    // ---- viscousFluxScheme::addFluxTerms, LaxFriedrichJacobian branch (viscousFluxScheme.C:220-246)
    if (c.mu > 0 && !c.sch.viscous_full_jacobian) {
        size_t n = (size_t)m.N + m.NB;
        vecd muEff(n, c.mu), alphaEff(n, c.gamma * (c.mu / c.Pr)), muf, alf, rhof, half(m.FT, 0.0);
        if (!c.muEffField.empty()) { muEff = c.muEffField; alphaEff = c.alphaEffField; syncCoupled(c, muEff, 1); syncCoupled(c, alphaEff, 1); }
        interpolateLinear(c, muEff, muf);
        interpolateLinear(c, alphaEff, alf);
        interpolateLinear(c, c.rho, rhof);
        for (int f = 0; f < m.FT; f++) if (faceActive(m, f, em)) half[f] = 0.5 * ((muf[f] + alf[f]) / rhof[f]);
        subtractLaplacianOne(m, dContByRho, half, {0});
        subtractLaplacianOne(m, dMomByRhoU, half, {0, 4, 8});
        subtractLaplacianOne(m, dEnergyByRhoE, half, {0});
    }
    // ---- LaxFriedrichJacobian false: viscousFluxScheme::addFluxTerms :248-261 + addBoundaryTerms :120-215
    if (c.mu > 0 && c.sch.viscous_full_jacobian) fullViscousJacobian(c);
    // addMRFSource, source part (convectiveFluxScheme.C:125-128): rhoU = rho*U, source -= (Omega ^ rhoU)*V.  The reference
    // applies it once to the fresh eqSystem of the iteration; the flag keeps a repeated assemble() from doing it twice.
    if (!c.mrfOmega.empty() && !c.srcMrfApplied) {
        for (int i = 0; i < m.N; i++) {
            const V3 om = {c.mrfOmega[3 * (size_t)i], c.mrfOmega[3 * (size_t)i + 1], c.mrfOmega[3 * (size_t)i + 2]};
            const V3 ru = c.rho[i] * V3{c.U[3 * (size_t)i], c.U[3 * (size_t)i + 1], c.U[3 * (size_t)i + 2]};
            const V3 cr = {om.y * ru.z - om.z * ru.y, om.z * ru.x - om.x * ru.z, om.x * ru.y - om.y * ru.x};
            c.srcRhoU[3 * (size_t)i] -= cr.x * m.V[i];
            c.srcRhoU[3 * (size_t)i + 1] -= cr.y * m.V[i];
            c.srcRhoU[3 * (size_t)i + 2] -= cr.z * m.V[i];
        }
        c.srcMrfApplied = true;
    }
    // sources (residualsUpdate.H:81-83)
    dContByRho.source = c.srcRho;
    dMomByRhoU.source = c.srcRhoU;
    dEnergyByRhoE.source = c.srcRhoE;
    c.matrixSet = true;
}

// ---------------------------------------------------------------------------------------------- update
// boundLocalTimeStep.H (executed before updateFields, i.e. on the not-yet-updated state — outerLoop.H:99)
void boundLocalTimeStep(Ctx& c)
{
    if (!(c.sch.local_timestepping && c.sch.local_timestepping_bounding)) return;
    const Mesh& m = c.m;
    const double lb = c.sch.local_timestepping_lower_bound;
    vecd rhoMin(m.N), eMin(m.N), eTemp(m.N), factor(m.N, 1.0);
    std::vector<char> bad(m.N);
    for (int i = 0; i < m.N; i++) {
        rhoMin[i] = lb * c.rhoPrev[i];
        V3 up = V3{c.rhoUPrev[3 * (size_t)i], c.rhoUPrev[3 * (size_t)i + 1], c.rhoUPrev[3 * (size_t)i + 2]} / c.rhoPrev[i];
        eMin[i] = lb * (c.rhoEPrev[i] / c.rhoPrev[i] - 0.5 * magSqr(up));
        V3 u = V3{c.rhoU[3 * (size_t)i], c.rhoU[3 * (size_t)i + 1], c.rhoU[3 * (size_t)i + 2]} / c.rho[i];
        eTemp[i] = c.rhoE[i] / c.rho[i] - 0.5 * magSqr(u);
        bad[i] = (c.rho[i] < rhoMin[i]) || (eTemp[i] < eMin[i]) || (eTemp[i] < SMALL);
    }
    for (int f = 0; f < m.F; f++) {
        int own = m.owner[f], nei = m.neighbour[f];
        if (bad[own]) { factor[own] = std::min(0.5, factor[own]); factor[nei] = std::min(0.75, factor[nei]); }
        if (bad[nei]) { factor[nei] = std::min(0.5, factor[nei]); factor[own] = std::min(0.75, factor[own]); }
    }
    vecd badv((size_t)m.N + m.NB, 0.0);
    for (int i = 0; i < m.N; i++) badv[i] = bad[i];
    syncCoupled(c, badv, 1);
    for (auto& p : m.patches)
        if (m.coupled(p))
            for (int f = p.start; f < p.start + p.size; f++) {
                int o = m.owner[f];
                if (bad[o]) factor[o] = std::min(0.5, factor[o]);
                if (badv[m.N + f - m.F] != 0.0) factor[o] = std::min(0.75, factor[o]);
            }
    for (int i = 0; i < m.N; i++) c.pseudoCoField[i] *= factor[i];
}

// updateFields.H
void updateFields(Ctx& c)
{
    const Mesh& m = c.m;
    bool boundLow = false, boundHigh = false;
    const double eBoundMin = c.Cv * c.sch.T_min, eBoundMax = c.Cv * c.sch.T_max;
    for (int i = 0; i < m.N; i++) {
        c.rho[i] += c.dRho[i];
        for (int d = 0; d < 3; d++) c.rhoU[3 * (size_t)i + d] += c.dRhoU[3 * (size_t)i + d];
        c.rhoE[i] += c.dRhoE[i];
    }
    // bound(rho, rhoMin): only acts when rho < rhoMin somewhere (rhoMin defaults to -GREAT)
    if (c.sch.rho_min > -GREAT)
        for (int i = 0; i < m.N; i++) c.rho[i] = std::max(c.rho[i], c.sch.rho_min);
    for (int i = 0; i < m.N; i++) {
        for (int d = 0; d < 3; d++) c.U[3 * (size_t)i + d] = c.rhoU[3 * (size_t)i + d] / c.rho[i];
        const double* u = &c.U[3 * (size_t)i];
        c.e[i] = c.rhoE[i] / c.rho[i] - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        if (c.e[i] - eBoundMin < 0) boundLow = true;
    }
    boundLow = c.comm->max(boundLow ? 1.0 : 0.0) > 0.5;
    if (boundLow) for (int i = 0; i < m.N; i++) c.e[i] = std::max(c.e[i], eBoundMin);
    if (c.sch.T_max < GREAT) {
        for (int i = 0; i < m.N; i++) if (c.e[i] - eBoundMax >= 0) boundHigh = true;
        boundHigh = c.comm->max(boundHigh ? 1.0 : 0.0) > 0.5;
        if (boundHigh) for (int i = 0; i < m.N; i++) c.e[i] = std::min(c.e[i], eBoundMax);
    }
    for (int i = 0; i < m.N; i++) {
        c.T[i] = c.e[i] / c.Cv;           // thermo.correct(): THE(e, p, T0) for hConst, Tref = 0
        c.psi[i] = 1.0 / (c.R * c.T[i]);  // perfectGas::psi
        c.p[i] = c.rho[i] / c.psi[i];
        const double* u = &c.U[3 * (size_t)i];
        for (int d = 0; d < 3; d++) c.rhoU[3 * (size_t)i + d] = c.rho[i] * u[d];
        c.rhoE[i] = c.rho[i] * (c.e[i] + 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]));
    }
    correctBoundary(c);
}

void newTimeStep(Ctx& c)
{
    const Mesh& m = c.m;
    c.rho00 = c.rho0; c.rhoU00 = c.rhoU0; c.rhoE00 = c.rhoE0;
    c.rho0.assign(c.rho.begin(), c.rho.begin() + m.N);
    c.rhoU0.assign(c.rhoU.begin(), c.rhoU.begin() + 3 * (size_t)m.N);
    c.rhoE0.assign(c.rhoE.begin(), c.rhoE.begin() + m.N);
    c.timeIndex++;
    // beginTimeStep.H:1-6: transient runs clear the SER residual memory every time step
    c.haveInitRes = c.havePrevRes = false;
    c.firstIter = true;
}

}  // namespace orc
