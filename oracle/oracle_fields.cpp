// oracle_fields.cpp — mesh setup, OpenFOAM field operators, thermo and boundary conditions of the oracle.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED (OpenFOAM arithmetic restated).
//
// Restates the OpenFOAM-v2112 pieces the reference path calls into:
//   gaussGrad::gradf, surfaceInterpolationScheme::interpolate, limitedSurfaceInterpolationScheme::weights,
//   NVDTVD::r, vanLeer / Minmod limiters, fvc::surfaceIntegrate (fvc::div), hePsiThermo/hConst/perfectGas,
//   fvPatchField::evaluate + valueInternalCoeffs for the BC set of the five tutorials.
#include "oracle_internal.hpp"

namespace orc {

// ------------------------------------------------------------------------------------------------ mesh
void meshFinalize(Ctx& c)
{
    Mesh& m = c.m;
    m.NB = m.FT - m.F;
    // primitiveMesh::calcCells ordering (used by lusgs.C:146 mesh.cells()[celli])
    veci cnt(m.N, 0);
    for (int f = 0; f < m.FT; f++) cnt[m.owner[f]]++;
    for (int f = 0; f < m.F; f++) cnt[m.neighbour[f]]++;
    m.cellFaceStart.assign(m.N + 1, 0);
    for (int i = 0; i < m.N; i++) m.cellFaceStart[i + 1] = m.cellFaceStart[i] + cnt[i];
    m.cellFaces.resize(m.cellFaceStart[m.N]);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int f = 0; f < m.FT; f++) { int o = m.owner[f]; m.cellFaces[m.cellFaceStart[o] + cnt[o]++] = f; }
    for (int f = 0; f < m.F; f++) { int n = m.neighbour[f]; m.cellFaces[m.cellFaceStart[n] + cnt[n]++] = f; }
    // coupled patch delta = own delta - neighbour delta (coupledFvPatch::delta / cyclicFvPatch::delta)
    m.dCoupled.assign(3 * (size_t)m.NB, 0.0);
    vecd own(3 * (size_t)m.NB, 0.0), nbr(3 * (size_t)m.NB, 0.0);
    for (auto& p : m.patches) {
        if (!m.coupled(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            int b = f - m.F, o = m.owner[f];
            for (int d = 0; d < 3; d++) own[3 * b + d] = m.Cf[3 * f + d] - m.C[3 * o + d];
        }
    }
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        auto& p = m.patches[pi];
        if (p.kind == ICSB200_CYCLIC) {
            auto& q = m.patches[p.nbrPatch];
            for (int i = 0; i < p.size; i++)
                {   // cyclicFvPatch::delta(): patchD - transform(forwardT, nbrPatchD)
                    const double* s3 = &own[3 * (size_t)(q.start + i - m.F)];
                    double* d3 = &nbr[3 * (size_t)(p.start + i - m.F)];
                    if (p.rotational) {
                        const double* T = p.forwardT;
                        d3[0] = T[0] * s3[0] + T[1] * s3[1] + T[2] * s3[2];
                        d3[1] = T[3] * s3[0] + T[4] * s3[1] + T[5] * s3[2];
                        d3[2] = T[6] * s3[0] + T[7] * s3[1] + T[8] * s3[2];
                    } else
                        for (int d = 0; d < 3; d++) d3[d] = s3[d];
                }
        } else if (p.kind == ICSB200_PROCESSOR) {
            c.comm->exchange(p.nbrRank, &own[3 * (p.start - m.F)], &nbr[3 * (p.start - m.F)], 3 * p.size);
        } else if (p.kind == ICSB200_CYCLICAMI) {
            // cyclicAMIFvPatch::delta: patchD - interpolate(nbrPatch.coupledFvPatch::delta())
            auto& q = m.patches[p.nbrPatch];
            for (int i = 0; i < p.size; i++) {
                double* d3 = &nbr[3 * (size_t)(p.start + i - m.F)];
                for (int d = 0; d < 3; d++) {
                    double acc = 0.0;
                    for (int k = p.amiStart[i]; k < p.amiStart[i + 1]; k++) acc += p.amiWeight[k] * own[3 * (q.start + p.amiFace[k] - m.F) + d];
                    d3[d] = acc;
                }
                if (p.rotational) {   // ... - transform(forwardT, interpolated neighbour delta)
                    const double* T = p.forwardT;
                    const double v0 = d3[0], v1 = d3[1], v2 = d3[2];
                    d3[0] = T[0] * v0 + T[1] * v1 + T[2] * v2;
                    d3[1] = T[3] * v0 + T[4] * v1 + T[5] * v2;
                    d3[2] = T[6] * v0 + T[7] * v1 + T[8] * v2;
                }
            }
        }
    }
    for (auto& p : m.patches)
        if (m.coupled(p))
            for (int b = p.start - m.F; b < p.start - m.F + p.size; b++)
                for (int d = 0; d < 3; d++) m.dCoupled[3 * b + d] = own[3 * b + d] - nbr[3 * b + d];
    c.bc.assign(m.patches.size(), {});
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        int k = m.coupled(m.patches[pi]) ? ICSB200_BC_COUPLED : m.empty(m.patches[pi]) ? ICSB200_BC_EMPTY : ICSB200_BC_ZEROGRADIENT;
        for (int fld = 0; fld < 3; fld++) c.bc[pi][fld].kind = k;
    }
}

// fill boundary slots of coupled patches with patchNeighbourField (nc doubles per value)
void syncCoupled(Ctx& c, vecd& vf, int nc)
{
    Mesh& m = c.m;
    for (auto& p : m.patches) {
        if (p.kind == ICSB200_CYCLIC) {
            // cyclicFvPatchField::patchNeighbourField (originalOFFiles/constraintFvPatchFields/cyclic/cyclicFvPatchField.C:130-190):
            // transform(forwardT, neighbour cell value) when the pair is rotational — vectors are rotated, scalars unchanged;
            // the scalar fields "U.component(i)" take component i of the rotated cell velocity, which is what the callers
            // get by extracting the components AFTER this call on U itself
            auto& q = m.patches[p.nbrPatch];
            const double* T = p.forwardT;
            for (int i = 0; i < p.size; i++) {
                int nb = m.owner[q.start + i];
                double* dst = &vf[(size_t)nc * (m.N + p.start - m.F + i)];
                const double* src = &vf[(size_t)nc * nb];
                if (p.rotational && nc == 3) {
                    const double v0 = src[0], v1 = src[1], v2 = src[2];
                    dst[0] = T[0] * v0 + T[1] * v1 + T[2] * v2;
                    dst[1] = T[3] * v0 + T[4] * v1 + T[5] * v2;
                    dst[2] = T[6] * v0 + T[7] * v1 + T[8] * v2;
                } else
                    for (int d = 0; d < nc; d++) dst[d] = src[d];
            }
        } else if (p.kind == ICSB200_CYCLICAMI) {
            // cyclicAMIFvPatchField::patchNeighbourField = AMI.interpolate(neighbour cell values): result = 0; result += w*phi
            auto& q = m.patches[p.nbrPatch];
            for (int i = 0; i < p.size; i++)
            {
                double* dst = &vf[(size_t)nc * (m.N + p.start - m.F + i)];
                for (int d = 0; d < nc; d++) {
                    double acc = 0.0;
                    for (int k = p.amiStart[i]; k < p.amiStart[i + 1]; k++) acc += p.amiWeight[k] * vf[(size_t)nc * m.owner[q.start + p.amiFace[k]] + d];
                    dst[d] = acc;
                }
                if (p.rotational && nc == 3) {   // cyclicAMIFvPatchField.C:171-203: transform(forwardT, interpolated value)
                    const double* T = p.forwardT;
                    const double v0 = dst[0], v1 = dst[1], v2 = dst[2];
                    dst[0] = T[0] * v0 + T[1] * v1 + T[2] * v2;
                    dst[1] = T[3] * v0 + T[4] * v1 + T[5] * v2;
                    dst[2] = T[6] * v0 + T[7] * v1 + T[8] * v2;
                }
            }
        } else if (p.kind == ICSB200_PROCESSOR) {
            vecd send((size_t)nc * p.size), recv((size_t)nc * p.size);
            for (int i = 0; i < p.size; i++)
                for (int d = 0; d < nc; d++) send[(size_t)nc * i + d] = vf[(size_t)nc * m.owner[p.start + i] + d];
            c.comm->exchange(p.nbrRank, send.data(), recv.data(), nc * p.size);
            for (int i = 0; i < p.size; i++)
                for (int d = 0; d < nc; d++) vf[(size_t)nc * (m.N + p.start - m.F + i) + d] = recv[(size_t)nc * i + d];
        }
    }
}

// phaseLagCyclicFvPatchField::patchNeighbourField / phaseLaggedField (phaseLagCyclicFvPatchField.C:160-398): for the fields
// the reference lists by name (U*, p, rho, E, H, c, gamma) the neighbour value of time instance K is
//   phaseLag = Zero; phaseLag += D_pl[K][J] * (field of instance J at the neighbour cell), J ascending; then transform(forwardT, .)
// — every other field (gradients, solver operands, T, ...) goes through the plain cyclic path of syncCoupled.
static double lagCellValue(const Ctx& s, int tag, int cell, int d)
{
    const double* u = &s.U[3 * (size_t)cell];
    switch (tag) {
        case LAG_P: return s.p[cell];
        case LAG_U: return u[d];
        case LAG_RHO: return s.rho[cell];
        default: break;
    }
    const double he = s.Cv * s.T[cell];
    const double E = he + 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    if (tag == LAG_E) return E;
    const double H = std::max(E, SMALL) + std::max(s.p[cell] / s.rho[cell], SMALL);
    if (tag == LAG_H) return H;
    if (tag == LAG_C2) return std::sqrt(2.0 * (s.gamma - 1.0) / (s.gamma + 1.0) * H);
    double cc = std::sqrt(s.gamma / s.psi[cell]);
    if (tag == LAG_C0) cc = std::max(cc, VSMALL);
    return cc;
}

void applyPhaseLag(Ctx& c, vecd& vf, int nc, int tag)
{
    if (c.lagRow.empty()) return;
    Mesh& m = c.m;
    for (auto& kv : c.lagRow) {
        const Patch& p = m.patches[kv.first];
        const Patch& q = m.patches[p.nbrPatch];
        const vecd& w = kv.second;
        for (int i = 0; i < p.size; i++) {
            const int nb = m.owner[q.start + i];
            double* dst = &vf[(size_t)nc * (m.N + p.start - m.F + i)];
            for (int d = 0; d < nc; d++) {
                double acc = 0.0;
                for (size_t J = 0; J < c.hbSiblings.size(); J++) acc += w[J] * lagCellValue(*c.hbSiblings[J], tag, nb, d);
                dst[d] = acc;
            }
            if (p.rotational && nc == 3) {
                const double* T = p.forwardT;
                const double v0 = dst[0], v1 = dst[1], v2 = dst[2];
                dst[0] = T[0] * v0 + T[1] * v1 + T[2] * v2;
                dst[1] = T[3] * v0 + T[4] * v1 + T[5] * v2;
                dst[2] = T[6] * v0 + T[7] * v1 + T[8] * v2;
            }
        }
    }
}

// the coupled-patch part of correctBoundary alone: patchNeighbourField caches of p, U, T, e, psi, rho, rhoU, rhoE.  With
// phase-lag patches the neighbour values mix all time instances, so they are refreshed once every instance has been updated
// (the reference evaluates patchNeighbourField() live whenever a face value is needed)
void resyncCoupledState(Ctx& c)
{
    Mesh& m = c.m;
    syncCoupled(c, c.p, 1); applyPhaseLag(c, c.p, 1, LAG_P);
    syncCoupled(c, c.U, 3); applyPhaseLag(c, c.U, 3, LAG_U);
    syncCoupled(c, c.T, 1);
    for (auto& p : m.patches) {
        if (!m.coupled(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            int s = m.N + f - m.F;
            c.e[s] = c.Cv * c.T[s];
            c.psi[s] = 1.0 / (c.R * c.T[s]);
        }
    }
    syncCoupled(c, c.rho, 1); applyPhaseLag(c, c.rho, 1, LAG_RHO);
    syncCoupled(c, c.rhoU, 3);
    syncCoupled(c, c.rhoE, 1);
}

// ------------------------------------------------------------------------------------------------ operators
// surfaceInterpolationScheme::interpolate with linear weights: sf = w*(vfP - vfN) + vfN; boundary: patch value,
// coupled: w*P + (1-w)*N
void interpolateLinear(const Ctx& c, const vecd& vf, vecd& sf)
{
    const Mesh& m = c.m;
    sf.assign(m.FT, 0.0);
    for (int f = 0; f < m.F; f++) {
        double P = vf[m.owner[f]], Nn = vf[m.neighbour[f]];
        sf[f] = m.w[f] * (P - Nn) + Nn;
    }
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            double b = vf[m.N + f - m.F];
            if (m.coupled(p)) sf[f] = m.w[f] * vf[m.owner[f]] + (1.0 - m.w[f]) * b;
            else sf[f] = b;
        }
    }
}

// gaussGrad<scalar>::gradf (Gauss linear) — grad holds 3*(N+NB); boundary slots only filled for coupled patches
void gradGauss(Ctx& c, const vecd& vf, vecd& grad)
{
    const Mesh& m = c.m;
    vecd ssf;
    interpolateLinear(c, vf, ssf);
    grad.assign(3 * (size_t)(m.N + m.NB), 0.0);
    for (int f = 0; f < m.F; f++) {
        int o = m.owner[f], n = m.neighbour[f];
        for (int d = 0; d < 3; d++) {
            double Sfssf = m.Sf[3 * f + d] * ssf[f];
            grad[3 * o + d] += Sfssf;
            grad[3 * n + d] -= Sfssf;
        }
    }
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            int o = m.owner[f];
            for (int d = 0; d < 3; d++) grad[3 * o + d] += m.Sf[3 * f + d] * ssf[f];
        }
    }
    for (int i = 0; i < m.N; i++)
        for (int d = 0; d < 3; d++) grad[3 * i + d] /= m.V[i];
    syncCoupled(c, grad, 3);
}

// NVDTVD::r
static inline double nvdR(double faceFlux, double phiP, double phiN, const double* gradcP, const double* gradcN, const double* d)
{
    double gradf = phiN - phiP;
    double gradcf;
    if (faceFlux > 0) gradcf = d[0] * gradcP[0] + d[1] * gradcP[1] + d[2] * gradcP[2];
    else gradcf = d[0] * gradcN[0] + d[1] * gradcN[1] + d[2] * gradcN[2];
    if (std::fabs(gradcf) >= 1000 * std::fabs(gradf)) return 2 * 1000 * sign(gradcf) * sign(gradf) - 1;
    return 2 * (gradcf / gradf) - 1;
}

static inline double limiterValue(int lim, double faceFlux, double phiP, double phiN, const double* gP, const double* gN, const double* d)
{
    switch (lim) {
        case ICSB200_LIM_VANLEER: { double r = nvdR(faceFlux, phiP, phiN, gP, gN, d); return (r + std::fabs(r)) / (1 + std::fabs(r)); }
        case ICSB200_LIM_MINMOD: { double r = nvdR(faceFlux, phiP, phiN, gP, gN, d); return std::max(std::min(r, 1.0), 0.0); }
        case ICSB200_LIM_LINEAR: return 1.0;
        default: return 0.0;  // upwind
    }
}

// fvc::interpolate(vf, dir, "reconstruct(..)") for dir = pos_ (+1) and neg_ (-1) at once
void interpolateLimitedLR(Ctx& c, const vecd& vf, int lim, vecd& sfL, vecd& sfR)
{
    const Mesh& m = c.m;
    vecd grad;
    if (lim == ICSB200_LIM_VANLEER || lim == ICSB200_LIM_MINMOD) gradGauss(c, vf, grad);
    else grad.assign(3 * (size_t)(m.N + m.NB), 0.0);
    sfL.assign(m.FT, 0.0);
    sfR.assign(m.FT, 0.0);
    for (int f = 0; f < m.F; f++) {
        int P = m.owner[f], Nn = m.neighbour[f];
        double d[3] = {m.C[3 * Nn] - m.C[3 * P], m.C[3 * Nn + 1] - m.C[3 * P + 1], m.C[3 * Nn + 2] - m.C[3 * P + 2]};
        double phiP = vf[P], phiN = vf[Nn];
        double limL = limiterValue(lim, 1.0, phiP, phiN, &grad[3 * P], &grad[3 * Nn], d);
        double limR = limiterValue(lim, -1.0, phiP, phiN, &grad[3 * P], &grad[3 * Nn], d);
        double wL = limL * m.w[f] + (1.0 - limL) * pos0(1.0);
        double wR = limR * m.w[f] + (1.0 - limR) * pos0(-1.0);
        sfL[f] = wL * (phiP - phiN) + phiN;
        sfR[f] = wR * (phiP - phiN) + phiN;
    }
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            int b = f - m.F;
            if (m.coupled(p)) {
                int P = m.owner[f];
                double phiP = vf[P], phiN = vf[m.N + b];
                const double* d = &m.dCoupled[3 * b];
                double limL = limiterValue(lim, 1.0, phiP, phiN, &grad[3 * P], &grad[3 * (size_t)(m.N + b)], d);
                double limR = limiterValue(lim, -1.0, phiP, phiN, &grad[3 * P], &grad[3 * (size_t)(m.N + b)], d);
                double wL = limL * m.w[f] + (1.0 - limL) * pos0(1.0);
                double wR = limR * m.w[f] + (1.0 - limR) * pos0(-1.0);
                sfL[f] = wL * phiP + (1.0 - wL) * phiN;
                sfR[f] = wR * phiP + (1.0 - wR) * phiN;
            } else {
                sfL[f] = vf[m.N + b];
                sfR[f] = vf[m.N + b];
            }
        }
    }
}

// fvc::surfaceIntegrate: div[nc per cell] of a surface field ssf[nc per face]
void surfaceIntegrate(const Ctx& c, const vecd& ssf, int nc, vecd& div)
{
    const Mesh& m = c.m;
    div.assign((size_t)nc * m.N, 0.0);
    for (int f = 0; f < m.F; f++)
        for (int d = 0; d < nc; d++) {
            div[(size_t)nc * m.owner[f] + d] += ssf[(size_t)nc * f + d];
            div[(size_t)nc * m.neighbour[f] + d] -= ssf[(size_t)nc * f + d];
        }
    for (auto& p : m.patches) {
        if (m.empty(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++)
            for (int d = 0; d < nc; d++) div[(size_t)nc * m.owner[f] + d] += ssf[(size_t)nc * f + d];
    }
    for (int i = 0; i < m.N; i++)
        for (int d = 0; d < nc; d++) div[(size_t)nc * i + d] /= m.V[i];
}

// ------------------------------------------------------------------------------------------------ BCs
// patch normal nf = Sf/magSf
static inline void nHat(const Mesh& m, int f, double n[3])
{
    for (int d = 0; d < 3; d++) n[d] = m.Sf[3 * f + d] / m.magSf[f];
}

// basicSymmetryFvPatchField<vector>::evaluate: (vP + transform(I - 2.0*sqr(nHat), vP))/2.0
static inline void symmetryVector(const double n[3], const double vP[3], double out[3])
{
    double xx = 1.0 - 2.0 * (n[0] * n[0]), xy = 0.0 - 2.0 * (n[0] * n[1]), xz = 0.0 - 2.0 * (n[0] * n[2]);
    double yy = 1.0 - 2.0 * (n[1] * n[1]), yz = 0.0 - 2.0 * (n[1] * n[2]), zz = 1.0 - 2.0 * (n[2] * n[2]);
    double t[3] = {xx * vP[0] + xy * vP[1] + xz * vP[2], xy * vP[0] + yy * vP[1] + yz * vP[2], xz * vP[0] + yz * vP[1] + zz * vP[2]};
    for (int d = 0; d < 3; d++) out[d] = (vP[d] + t[d]) / 2.0;
}

// evaluate p on one patch (p.correctBoundaryConditions(), updateFields.H:80)
void evalP(Ctx& c, int pi)
{
    Mesh& m = c.m;
    auto& p = m.patches[pi];
    const BC& bc = c.bc[pi][ICSB200_FIELD_P];
    for (int f = p.start; f < p.start + p.size; f++) {
        int b = f - m.F, o = m.owner[f], s = m.N + b;
        switch (bc.kind) {
            case ICSB200_BC_ZEROGRADIENT:
            case ICSB200_BC_SLIP: c.p[s] = c.p[o]; break;
            case ICSB200_BC_FIXEDVALUE: c.p[s] = bc.P(f - p.start)[0]; break;
            case ICSB200_BC_INLETOUTLET: {
                double vfrac = 1.0 - pos0(c.phi[f]);
                c.p[s] = vfrac * bc.P(f - p.start)[0] + (1.0 - vfrac) * (c.p[o] + 0.0);
                break;
            }
            case ICSB200_BC_TOTALPRESSURE: {
                // totalPressureFvPatchScalarField::updateCoeffs, high-speed compressible branch
                double p0 = bc.P(f - p.start)[0], g = bc.P(f - p.start)[1];
                const double* Up = &c.U[3 * (size_t)s];
                double magSqrUp = Up[0] * Up[0] + Up[1] * Up[1] + Up[2] * Up[2];
                double psip = c.psi[s];
                if (g > 1) {
                    double gM1ByG = (g - 1) / g;
                    c.p[s] = p0 / std::pow(1.0 + 0.5 * psip * gM1ByG * (1.0 - pos0(c.phi[f])) * magSqrUp, 1 / gM1ByG);
                } else {
                    c.p[s] = p0 / (1.0 + 0.5 * psip * (1.0 - pos0(c.phi[f])) * magSqrUp);
                }
                break;
            }
            case ICSB200_BC_FREESTREAMPRESSURE: {
                // freestreamPressureFvPatchScalarField::updateCoeffs (subsonic branch) + mixed evaluate
                const double* Ui = &bc.P(f - p.start)[1];
                double magUp = std::sqrt(Ui[0] * Ui[0] + Ui[1] * Ui[1] + Ui[2] * Ui[2]);
                double n[3];
                nHat(m, f, n);
                double vfrac = 0.5;
                if (magUp > VSMALL) vfrac = 0.5 + 0.5 * (Ui[0] * n[0] + Ui[1] * n[1] + Ui[2] * n[2]) / magUp;
                c.p[s] = vfrac * bc.P(f - p.start)[0] + (1.0 - vfrac) * (c.p[o] + 0.0);
                break;
            }
            default: break;
        }
    }
}

void evalU(Ctx& c, int pi)
{
    Mesh& m = c.m;
    auto& p = m.patches[pi];
    const BC& bc = c.bc[pi][ICSB200_FIELD_U];
    for (int f = p.start; f < p.start + p.size; f++) {
        int b = f - m.F, o = m.owner[f], s = m.N + b;
        double* Ub = &c.U[3 * (size_t)s];
        const double* UP = &c.U[3 * (size_t)o];
        switch (bc.kind) {
            case ICSB200_BC_ZEROGRADIENT: for (int d = 0; d < 3; d++) Ub[d] = UP[d]; break;
            case ICSB200_BC_FIXEDVALUE: for (int d = 0; d < 3; d++) Ub[d] = bc.P(f - p.start)[d]; break;
            case ICSB200_BC_SLIP: { double n[3]; nHat(m, f, n); symmetryVector(n, UP, Ub); break; }
            case ICSB200_BC_INLETOUTLET: {
                double vfrac = 1.0 - pos0(c.phi[f]);
                for (int d = 0; d < 3; d++) Ub[d] = vfrac * bc.P(f - p.start)[d] + (1.0 - vfrac) * (UP[d] + 0.0);
                break;
            }
            case ICSB200_BC_PRESSUREINLETOUTLETVELOCITY: {
                // valueFraction = neg(phip)*(I - sqr(nf)); refValue = tv - n(n & tv); directionMixed::evaluate
                double n[3];
                nHat(m, f, n);
                double sgn = neg(c.phi[f]);
                double vf[6] = {sgn * (1.0 - n[0] * n[0]), sgn * (0.0 - n[0] * n[1]), sgn * (0.0 - n[0] * n[2]),
                                sgn * (1.0 - n[1] * n[1]), sgn * (0.0 - n[1] * n[2]), sgn * (1.0 - n[2] * n[2])};
                const double* tv = bc.P(f - p.start);
                double ntv = n[0] * tv[0] + n[1] * tv[1] + n[2] * tv[2];
                double ref[3] = {tv[0] - n[0] * ntv, tv[1] - n[1] * ntv, tv[2] - n[2] * ntv};
                double nv[3] = {vf[0] * ref[0] + vf[1] * ref[1] + vf[2] * ref[2], vf[1] * ref[0] + vf[3] * ref[1] + vf[4] * ref[2],
                                vf[2] * ref[0] + vf[4] * ref[1] + vf[5] * ref[2]};
                double g[3] = {UP[0] + 0.0, UP[1] + 0.0, UP[2] + 0.0};
                double iv[6] = {1.0 - vf[0], 0.0 - vf[1], 0.0 - vf[2], 1.0 - vf[3], 0.0 - vf[4], 1.0 - vf[5]};
                double tg[3] = {iv[0] * g[0] + iv[1] * g[1] + iv[2] * g[2], iv[1] * g[0] + iv[3] * g[1] + iv[4] * g[2],
                                iv[2] * g[0] + iv[4] * g[1] + iv[5] * g[2]};
                for (int d = 0; d < 3; d++) Ub[d] = nv[d] + tg[d];
                break;
            }
            default: break;
        }
    }
}

void evalT(Ctx& c, int pi)
{
    Mesh& m = c.m;
    auto& p = m.patches[pi];
    const BC& bc = c.bc[pi][ICSB200_FIELD_T];
    for (int f = p.start; f < p.start + p.size; f++) {
        int b = f - m.F, o = m.owner[f], s = m.N + b;
        switch (bc.kind) {
            case ICSB200_BC_ZEROGRADIENT:
            case ICSB200_BC_SLIP: c.T[s] = c.T[o]; break;
            case ICSB200_BC_FIXEDVALUE: c.T[s] = bc.P(f - p.start)[0]; break;
            case ICSB200_BC_INLETOUTLET: {
                double vfrac = 1.0 - pos0(c.phi[f]);
                c.T[s] = vfrac * bc.P(f - p.start)[0] + (1.0 - vfrac) * (c.T[o] + 0.0);
                break;
            }
            case ICSB200_BC_TOTALTEMPERATURE: {
                double T0 = bc.P(f - p.start)[0], g = bc.P(f - p.start)[1];
                double gM1ByG = (g - 1) / g;
                const double* Up = &c.U[3 * (size_t)s];
                double magSqrUp = Up[0] * Up[0] + Up[1] * Up[1] + Up[2] * Up[2];
                c.T[s] = T0 / (1.0 + 0.5 * c.psi[s] * gM1ByG * (1.0 - pos0(c.phi[f])) * magSqrUp);
                break;
            }
            default: break;
        }
    }
}

// does the T patch field fix its value?  (hePsiThermo::calculate boundary branch)
static bool fixesValueT(int kind) { return kind == ICSB200_BC_FIXEDVALUE || kind == ICSB200_BC_TOTALTEMPERATURE; }

// updateFields.H:80-104 — boundary refresh of p, U, T then thermo and conserved boundary values
void correctBoundary(Ctx& c)
{
    Mesh& m = c.m;
    for (size_t pi = 0; pi < m.patches.size(); pi++) if (!m.empty(m.patches[pi]) && !m.coupled(m.patches[pi])) evalP(c, (int)pi);
    for (size_t pi = 0; pi < m.patches.size(); pi++) if (!m.empty(m.patches[pi]) && !m.coupled(m.patches[pi])) evalU(c, (int)pi);
    for (size_t pi = 0; pi < m.patches.size(); pi++) if (!m.empty(m.patches[pi]) && !m.coupled(m.patches[pi])) evalT(c, (int)pi);
    c.vicP.assign(m.NB, 0.0); c.vicU.assign(3 * (size_t)m.NB, 0.0); c.vicT.assign(m.NB, 0.0);
    c.gicU.assign(3 * (size_t)m.NB, 0.0); c.gicT.assign(m.NB, 0.0);
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        auto& p = m.patches[pi];
        if (m.empty(p) || m.coupled(p)) continue;
        for (int f = p.start; f < p.start + p.size; f++) {
            int b = f - m.F;
            valueInternalCoeffs(c, (int)pi, f, c.vicP[b], &c.vicU[3 * (size_t)b], c.vicT[b]);
            gradientInternalCoeffs(c, (int)pi, f, &c.gicU[3 * (size_t)b], c.gicT[b]);
        }
    }
    syncCoupled(c, c.p, 1); applyPhaseLag(c, c.p, 1, LAG_P);
    syncCoupled(c, c.U, 3); applyPhaseLag(c, c.U, 3, LAG_U);
    syncCoupled(c, c.T, 1);
    for (size_t pi = 0; pi < m.patches.size(); pi++) {
        auto& p = m.patches[pi];
        if (m.empty(p)) continue;
        bool fixes = fixesValueT(c.bc[pi][ICSB200_FIELD_T].kind);
        for (int f = p.start; f < p.start + p.size; f++) {
            int s = m.N + f - m.F;
            if (!m.coupled(p)) {
                // T.boundaryFieldRef() == max(T.boundaryField(), TMin)  (updateFields.H:84; Q5: TMax uses min here)
                c.T[s] = std::max(c.T[s], c.sch.T_min);
                if (c.sch.T_max < GREAT) c.T[s] = std::min(c.T[s], c.sch.T_max);
            }
            c.e[s] = c.Cv * c.T[s];                   // e.boundaryFieldRef() == thermo.he(p, T)
            if (!fixes && !m.coupled(p)) c.T[s] = c.e[s] / c.Cv;  // thermo.correct(): T = THE(he, p, T0)
            c.psi[s] = 1.0 / (c.R * c.T[s]);
            if (m.coupled(p)) {
                // coupled patches carry the neighbour cell's conserved values
                continue;
            }
            c.rho[s] = c.psi[s] * c.p[s];
            const double* Ub = &c.U[3 * (size_t)s];
            for (int d = 0; d < 3; d++) c.rhoU[3 * (size_t)s + d] = c.rho[s] * Ub[d];
            c.rhoE[s] = c.rho[s] * (c.e[s] + 0.5 * (Ub[0] * Ub[0] + Ub[1] * Ub[1] + Ub[2] * Ub[2]));
        }
    }
    syncCoupled(c, c.rho, 1); applyPhaseLag(c, c.rho, 1, LAG_RHO);
    syncCoupled(c, c.rhoU, 3);
    syncCoupled(c, c.rhoE, 1);
}

// valueInternalCoeffs of the p / U / T patch fields (convectiveFluxScheme.C:67-78)
void valueInternalCoeffs(const Ctx& c, int pi, int f, double& pVIC, double uVIC[3], double& tVIC)
{
    const Mesh& m = c.m;
    auto vfracPhi = [&]() { return 1.0 - pos0(c.phi[f]); };
    const BC& bp = c.bc[pi][ICSB200_FIELD_P];
    switch (bp.kind) {
        case ICSB200_BC_ZEROGRADIENT: case ICSB200_BC_SLIP: pVIC = 1.0; break;
        case ICSB200_BC_INLETOUTLET: pVIC = 1.0 * (1.0 - vfracPhi()); break;
        case ICSB200_BC_FREESTREAMPRESSURE: {
            const double* Ui = &bp.P(f - m.patches[pi].start)[1];
            double magUp = std::sqrt(Ui[0] * Ui[0] + Ui[1] * Ui[1] + Ui[2] * Ui[2]);
            double n[3];
            nHat(m, f, n);
            double vfrac = 0.5;
            if (magUp > VSMALL) vfrac = 0.5 + 0.5 * (Ui[0] * n[0] + Ui[1] * n[1] + Ui[2] * n[2]) / magUp;
            pVIC = 1.0 * (1.0 - vfrac);
            break;
        }
        default: pVIC = 0.0; break;  // fixedValue, totalPressure
    }
    const BC& bu = c.bc[pi][ICSB200_FIELD_U];
    switch (bu.kind) {
        case ICSB200_BC_ZEROGRADIENT: uVIC[0] = uVIC[1] = uVIC[2] = 1.0; break;
        case ICSB200_BC_SLIP: { double n[3]; nHat(m, f, n); for (int d = 0; d < 3; d++) uVIC[d] = 1.0 - std::fabs(n[d]); break; }
        case ICSB200_BC_INLETOUTLET: for (int d = 0; d < 3; d++) uVIC[d] = 1.0 * (1.0 - vfracPhi()); break;
        case ICSB200_BC_PRESSUREINLETOUTLETVELOCITY: {
            double n[3];
            nHat(m, f, n);
            double sgn = neg(c.phi[f]);
            for (int d = 0; d < 3; d++) uVIC[d] = 1.0 - std::sqrt(std::fabs(sgn * (1.0 - n[d] * n[d])));
            break;
        }
        default: uVIC[0] = uVIC[1] = uVIC[2] = 0.0; break;
    }
    const BC& bt = c.bc[pi][ICSB200_FIELD_T];
    switch (bt.kind) {
        case ICSB200_BC_ZEROGRADIENT: case ICSB200_BC_SLIP: tVIC = 1.0; break;
        case ICSB200_BC_INLETOUTLET: tVIC = 1.0 * (1.0 - vfracPhi()); break;
        default: tVIC = 0.0; break;
    }
}

// gradientInternalCoeffs of the U / T patch fields (viscousFluxScheme.C:58-59): fixedValue -deltaCoeffs; zeroGradient 0;
// mixed -valueFraction deltaCoeffs; transform (basicSymmetry, directionMixed) -deltaCoeffs snGradTransformDiag(), with
// snGradTransformDiag = |nHat_d| for the symmetry family and sqrt(|valueFraction_dd|) for directionMixed; scalars on
// transform patches 0 (transformFvPatchScalarField)
void gradientInternalCoeffs(const Ctx& c, int pi, int f, double uGIC[3], double& tGIC)
{
    const Mesh& m = c.m;
    const double dc = m.deltaCoeffs[f];
    auto vfracPhi = [&]() { return 1.0 - pos0(c.phi[f]); };
    const BC& bu = c.bc[pi][ICSB200_FIELD_U];
    switch (bu.kind) {
        case ICSB200_BC_ZEROGRADIENT: uGIC[0] = uGIC[1] = uGIC[2] = 0.0; break;
        case ICSB200_BC_SLIP: { double n[3]; nHat(m, f, n); for (int d = 0; d < 3; d++) uGIC[d] = -dc * std::fabs(n[d]); break; }
        case ICSB200_BC_INLETOUTLET: for (int d = 0; d < 3; d++) uGIC[d] = -1.0 * vfracPhi() * dc; break;
        case ICSB200_BC_PRESSUREINLETOUTLETVELOCITY: {
            double n[3];
            nHat(m, f, n);
            double sgn = neg(c.phi[f]);
            for (int d = 0; d < 3; d++) uGIC[d] = -dc * std::sqrt(std::fabs(sgn * (1.0 - n[d] * n[d])));
            break;
        }
        default: uGIC[0] = uGIC[1] = uGIC[2] = -1.0 * dc; break;  // fixedValue family
    }
    const BC& bt = c.bc[pi][ICSB200_FIELD_T];
    switch (bt.kind) {
        case ICSB200_BC_ZEROGRADIENT: case ICSB200_BC_SLIP: tGIC = 0.0; break;
        case ICSB200_BC_INLETOUTLET: tGIC = -1.0 * vfracPhi() * dc; break;
        default: tGIC = -1.0 * dc; break;  // fixedValue, totalTemperature
    }
}

// ------------------------------------------------------------------------------------------------ state
// createFields.H:75-146 — conserved variables from p, U, T (cells), boundary init, phi = Sf & interpolate(rhoU)
void stateInit(Ctx& c, const double* p, const double* U, const double* T)
{
    Mesh& m = c.m;
    size_t n = (size_t)m.N + m.NB;
    c.p.assign(n, 0); c.T.assign(n, 0); c.e.assign(n, 0); c.psi.assign(n, 0); c.rho.assign(n, 0); c.rhoE.assign(n, 0);
    c.U.assign(3 * n, 0); c.rhoU.assign(3 * n, 0);
    for (int i = 0; i < m.N; i++) {
        c.p[i] = p[i]; c.T[i] = T[i];
        for (int d = 0; d < 3; d++) c.U[3 * (size_t)i + d] = U[3 * (size_t)i + d];
        c.e[i] = c.Cv * c.T[i];
        c.psi[i] = 1.0 / (c.R * c.T[i]);
        c.rho[i] = c.psi[i] * c.p[i];
        const double* Ui = &c.U[3 * (size_t)i];
        for (int d = 0; d < 3; d++) c.rhoU[3 * (size_t)i + d] = c.rho[i] * Ui[d];
        c.rhoE[i] = c.rho[i] * (c.e[i] + 0.5 * (Ui[0] * Ui[0] + Ui[1] * Ui[1] + Ui[2] * Ui[2]));
    }
    // provisional boundary state = adjacent cell (stands in for the 'value' entries OpenFOAM reads from 0/)
    c.phi.assign(m.FT, 0); c.phiUp.assign(3 * (size_t)m.FT, 0); c.phiEp.assign(m.FT, 0);
    for (auto& pa : m.patches) {
        if (m.empty(pa)) continue;
        for (int f = pa.start; f < pa.start + pa.size; f++) {
            int o = m.owner[f], s = m.N + f - m.F;
            c.p[s] = c.p[o]; c.T[s] = c.T[o]; c.psi[s] = c.psi[o]; c.e[s] = c.e[o]; c.rho[s] = c.rho[o]; c.rhoE[s] = c.rhoE[o];
            for (int d = 0; d < 3; d++) { c.U[3 * (size_t)s + d] = c.U[3 * (size_t)o + d]; c.rhoU[3 * (size_t)s + d] = c.rhoU[3 * (size_t)o + d]; }
            c.phi[f] = m.Sf[3 * f] * c.rhoU[3 * (size_t)o] + m.Sf[3 * f + 1] * c.rhoU[3 * (size_t)o + 1] + m.Sf[3 * f + 2] * c.rhoU[3 * (size_t)o + 2];
        }
    }
    correctBoundary(c);
    correctBoundary(c);
    // phi = mesh.Sf() & fvc::interpolate(rhoU)   (createFields.H:133-146)
    vecd comp((size_t)m.N + m.NB), sf;
    std::fill(c.phi.begin(), c.phi.end(), 0.0);
    vecd acc[3];
    for (int d = 0; d < 3; d++) {
        for (size_t i = 0; i < n; i++) comp[i] = c.rhoU[3 * i + d];
        interpolateLinear(c, comp, acc[d]);
    }
    for (int f = 0; f < m.FT; f++) c.phi[f] = m.Sf[3 * f] * acc[0][f] + m.Sf[3 * f + 1] * acc[1][f] + m.Sf[3 * f + 2] * acc[2][f];
    c.rho0 = vecd(c.rho.begin(), c.rho.begin() + m.N); c.rho00 = c.rho0;
    c.rhoU0 = vecd(c.rhoU.begin(), c.rhoU.begin() + 3 * (size_t)m.N); c.rhoU00 = c.rhoU0;
    c.rhoE0 = vecd(c.rhoE.begin(), c.rhoE.begin() + m.N); c.rhoE00 = c.rhoE0;
    c.timeIndex = 0;
    c.rPseudoDeltaT.assign(m.N, 0.0);
    c.pseudoCoField.assign(m.N, c.sch.pseudo_co_num);
    c.pseudoCoNum = c.sch.pseudo_co_num;
    c.ddtCoeff.assign(m.N, 0.0);
    c.haveInitRes = c.havePrevRes = false;
    c.firstIter = true;
    c.dRho.assign(m.N, 0); c.dRhoU.assign(3 * (size_t)m.N, 0); c.dRhoE.assign(m.N, 0);
    c.stateSet = true;
}

}  // namespace orc
