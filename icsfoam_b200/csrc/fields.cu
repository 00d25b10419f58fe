// fields.cu — state upload/download, boundary conditions, thermo/derived fields, field update.
//
// Device counterparts of: createFields.H:75-146 (state from p,U,T), updateFields.H:1-104 (W += dW, primitives,
// energy bounding, thermo.correct(), BC refresh), boundLocalTimeStep.H:1-98, the fvPatchField::evaluate /
// valueInternalCoeffs of the BC set used by the five tutorials (convectiveFluxScheme.C:67-78), and the
// per-cell derived fields every flux scheme builds before reconstructing (hllcFluxScheme.C:100-121,
// ausmPlusUpFluxScheme.C:105-121).  One thread per cell position / boundary face; SoA, coalesced.
#include "common.cuh"

namespace {

struct Thermo { double R, Cv, gamma; };

__device__ __forceinline__ double pos0(double s) { return s >= 0 ? 1.0 : 0.0; }
__device__ __forceinline__ double negf(double s) { return s < 0 ? 1.0 : 0.0; }

// createFields.H:75-131: e, psi, rho, rhoU, rhoE from p, U, T on cells
__global__ void k_state_init(int NP, const int* __restrict__ pos2cell, Thermo th, double* __restrict__ f, size_t NX)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    double T = f[Q_T * NX + p], pr = f[Q_P * NX + p];
    double ux = f[Q_UX * NX + p], uy = f[Q_UY * NX + p], uz = f[Q_UZ * NX + p];
    double e = th.Cv * T;
    double psi = 1.0 / (th.R * T);
    double rho = psi * pr;
    f[Q_PSI * NX + p] = psi;
    f[Q_RHO * NX + p] = rho;
    f[Q_W0 * NX + p] = rho;
    f[Q_W1 * NX + p] = rho * ux;
    f[Q_W2 * NX + p] = rho * uy;
    f[Q_W3 * NX + p] = rho * uz;
    f[Q_W4 * NX + p] = rho * (e + 0.5 * (ux * ux + uy * uy + uz * uz));
}

// provisional boundary state = adjacent cell; phi_b = Sf & rhoU_P (stands in for the 'value' entries of 0/*)
__global__ void k_boundary_provisional(int NB, int off, const int* __restrict__ bfOwnerPos, const int* __restrict__ bfPatch,
                                       const double* __restrict__ bfGeo, double* __restrict__ f, size_t NX, double* __restrict__ phiB)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB || bfPatch[b] < 0) return;
    int o = bfOwnerPos[b];
    size_t s = (size_t)off + b;
    for (int q : {Q_P, Q_T, Q_PSI, Q_RHO, Q_UX, Q_UY, Q_UZ}) f[q * NX + s] = f[q * NX + o];
    phiB[b] = bfGeo[b] * f[Q_W1 * NX + o] + bfGeo[NB + b] * f[Q_W2 * NX + o] + bfGeo[2 * (size_t)NB + b] * f[Q_W3 * NX + o];
}

// phi_b = Sf & (rho_b U_b)   (createFields.H:133-146 restricted to boundary faces)
__global__ void k_boundary_phi(int NB, int off, const int* __restrict__ bfPatch, const BCDev* __restrict__ bc, const double* __restrict__ bfGeo,
                               const double* __restrict__ f, size_t NX, double* __restrict__ phiB)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB || bfPatch[b] < 0) return;
    if (bc[bfPatch[b]].kind[0] == ICSB200_BC_COUPLED) return;
    size_t s = (size_t)off + b;
    double rho = f[Q_RHO * NX + s];
    double rux = rho * f[Q_UX * NX + s], ruy = rho * f[Q_UY * NX + s], ruz = rho * f[Q_UZ * NX + s];
    phiB[b] = bfGeo[b] * rux + bfGeo[NB + b] * ruy + bfGeo[2 * (size_t)NB + b] * ruz;
}

// p.correctBoundaryConditions(); U...; T...; clamp; e_b; thermo.correct(); rho_b  (updateFields.H:80-104)
// + valueInternalCoeffs frozen for addBoundaryTerms (convectiveFluxScheme.C:67-78)
__global__ void k_bc(int NB, int off, const int* __restrict__ bfOwnerPos, const int* __restrict__ bfPatch, const BCDev* __restrict__ bcs,
                     const double* __restrict__ bfGeo, const double* __restrict__ phiB, Thermo th, double TMin, double TMax,
                     double* __restrict__ f, size_t NX, double* __restrict__ vic)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    int pi = bfPatch[b];
    if (pi < 0) return;
    const BCDev& bc = bcs[pi];
    if (bc.kind[0] == ICSB200_BC_COUPLED) return;
    const int o = bfOwnerPos[b];
    const size_t s = (size_t)off + b;
    const double magSf = bfGeo[3 * (size_t)NB + b];
    const double n[3] = {bfGeo[b] / magSf, bfGeo[NB + b] / magSf, bfGeo[2 * (size_t)NB + b] / magSf};
    const double phi = phiB[b];
    const double pP = f[Q_P * NX + o], TP = f[Q_T * NX + o];
    const double UP[3] = {f[Q_UX * NX + o], f[Q_UY * NX + o], f[Q_UZ * NX + o]};
    const double psiOld = f[Q_PSI * NX + s];
    double Ub[3] = {f[Q_UX * NX + s], f[Q_UY * NX + s], f[Q_UZ * NX + s]};  // previous boundary velocity (totalPressure reads it)
    double pb = pP, Tb = TP, pVIC = 1.0, tVIC = 1.0, uVIC[3] = {1.0, 1.0, 1.0};
    // gradientInternalCoeffs = -deltaCoeffs * g (viscousFluxScheme.C:58-59): g = 1 fixedValue family, 0 zeroGradient, valueFraction for
    // mixed, snGradTransformDiag for the transform family (|nHat_d| symmetry, sqrt|valueFraction_dd| directionMixed; scalars 0)
    double gU[3] = {1.0, 1.0, 1.0}, gT = 1.0;
    const double vfracPhi = 1.0 - pos0(phi);
    // ---- p
    {
        const double* prm = bc.P(0, b);
        switch (bc.kind[0]) {
            case ICSB200_BC_ZEROGRADIENT:
            case ICSB200_BC_SLIP: pb = pP; pVIC = 1.0; break;
            case ICSB200_BC_FIXEDVALUE: pb = prm[0]; pVIC = 0.0; break;
            case ICSB200_BC_INLETOUTLET: pb = vfracPhi * prm[0] + (1.0 - vfracPhi) * (pP + 0.0); pVIC = 1.0 * (1.0 - vfracPhi); break;
            case ICSB200_BC_TOTALPRESSURE: {
                double p0 = prm[0], g = prm[1];
                double magSqrUp = Ub[0] * Ub[0] + Ub[1] * Ub[1] + Ub[2] * Ub[2];
                if (g > 1) {
                    double gM1ByG = (g - 1) / g;
                    pb = p0 / pow(1.0 + 0.5 * psiOld * gM1ByG * (1.0 - pos0(phi)) * magSqrUp, 1 / gM1ByG);
                } else {
                    pb = p0 / (1.0 + 0.5 * psiOld * (1.0 - pos0(phi)) * magSqrUp);
                }
                pVIC = 0.0;
                break;
            }
            case ICSB200_BC_FREESTREAMPRESSURE: {
                const double* Ui = &prm[1];
                double magUp = sqrt(Ui[0] * Ui[0] + Ui[1] * Ui[1] + Ui[2] * Ui[2]);
                double vfrac = 0.5;
                if (magUp > ICS_VSMALL) vfrac = 0.5 + 0.5 * (Ui[0] * n[0] + Ui[1] * n[1] + Ui[2] * n[2]) / magUp;
                pb = vfrac * prm[0] + (1.0 - vfrac) * (pP + 0.0);
                pVIC = 1.0 * (1.0 - vfrac);
                break;
            }
            default: pb = pP; pVIC = 0.0; break;
        }
    }
    // ---- U
    {
        const double* prm = bc.P(1, b);
        switch (bc.kind[1]) {
            case ICSB200_BC_ZEROGRADIENT: for (int d = 0; d < 3; d++) { Ub[d] = UP[d]; uVIC[d] = 1.0; gU[d] = 0.0; } break;
            case ICSB200_BC_FIXEDVALUE: for (int d = 0; d < 3; d++) { Ub[d] = prm[d]; uVIC[d] = 0.0; } break;
            case ICSB200_BC_SLIP: {
                // basicSymmetryFvPatchField<vector>::evaluate
                double xx = 1.0 - 2.0 * (n[0] * n[0]), xy = 0.0 - 2.0 * (n[0] * n[1]), xz = 0.0 - 2.0 * (n[0] * n[2]);
                double yy = 1.0 - 2.0 * (n[1] * n[1]), yz = 0.0 - 2.0 * (n[1] * n[2]), zz = 1.0 - 2.0 * (n[2] * n[2]);
                double t[3] = {xx * UP[0] + xy * UP[1] + xz * UP[2], xy * UP[0] + yy * UP[1] + yz * UP[2], xz * UP[0] + yz * UP[1] + zz * UP[2]};
                for (int d = 0; d < 3; d++) { Ub[d] = (UP[d] + t[d]) / 2.0; uVIC[d] = 1.0 - fabs(n[d]); gU[d] = fabs(n[d]); }
                break;
            }
            case ICSB200_BC_INLETOUTLET:
                for (int d = 0; d < 3; d++) { Ub[d] = vfracPhi * prm[d] + (1.0 - vfracPhi) * (UP[d] + 0.0); uVIC[d] = 1.0 * (1.0 - vfracPhi); gU[d] = vfracPhi; }
                break;
            case ICSB200_BC_PRESSUREINLETOUTLETVELOCITY: {
                double sgn = negf(phi);
                double vf[6] = {sgn * (1.0 - n[0] * n[0]), sgn * (0.0 - n[0] * n[1]), sgn * (0.0 - n[0] * n[2]),
                                sgn * (1.0 - n[1] * n[1]), sgn * (0.0 - n[1] * n[2]), sgn * (1.0 - n[2] * n[2])};
                double ntv = n[0] * prm[0] + n[1] * prm[1] + n[2] * prm[2];
                double ref[3] = {prm[0] - n[0] * ntv, prm[1] - n[1] * ntv, prm[2] - n[2] * ntv};
                double nv[3] = {vf[0] * ref[0] + vf[1] * ref[1] + vf[2] * ref[2], vf[1] * ref[0] + vf[3] * ref[1] + vf[4] * ref[2],
                                vf[2] * ref[0] + vf[4] * ref[1] + vf[5] * ref[2]};
                double g[3] = {UP[0] + 0.0, UP[1] + 0.0, UP[2] + 0.0};
                double iv[6] = {1.0 - vf[0], 0.0 - vf[1], 0.0 - vf[2], 1.0 - vf[3], 0.0 - vf[4], 1.0 - vf[5]};
                double tg[3] = {iv[0] * g[0] + iv[1] * g[1] + iv[2] * g[2], iv[1] * g[0] + iv[3] * g[1] + iv[4] * g[2],
                                iv[2] * g[0] + iv[4] * g[1] + iv[5] * g[2]};
                for (int d = 0; d < 3; d++) { Ub[d] = nv[d] + tg[d]; uVIC[d] = 1.0 - sqrt(fabs(sgn * (1.0 - n[d] * n[d]))); gU[d] = sqrt(fabs(sgn * (1.0 - n[d] * n[d]))); }
                break;
            }
            default: for (int d = 0; d < 3; d++) { Ub[d] = UP[d]; uVIC[d] = 0.0; } break;
        }
    }
    // ---- T
    bool fixesT = false;
    {
        const double* prm = bc.P(2, b);
        switch (bc.kind[2]) {
            case ICSB200_BC_ZEROGRADIENT:
            case ICSB200_BC_SLIP: Tb = TP; tVIC = 1.0; gT = 0.0; break;
            case ICSB200_BC_FIXEDVALUE: Tb = prm[0]; tVIC = 0.0; fixesT = true; break;
            case ICSB200_BC_INLETOUTLET: Tb = vfracPhi * prm[0] + (1.0 - vfracPhi) * (TP + 0.0); tVIC = 1.0 * (1.0 - vfracPhi); gT = vfracPhi; break;
            case ICSB200_BC_TOTALTEMPERATURE: {
                double T0 = prm[0], g = prm[1];
                double gM1ByG = (g - 1) / g;
                double magSqrUp = Ub[0] * Ub[0] + Ub[1] * Ub[1] + Ub[2] * Ub[2];
                Tb = T0 / (1.0 + 0.5 * psiOld * gM1ByG * (1.0 - pos0(phi)) * magSqrUp);
                tVIC = 0.0;
                fixesT = true;
                break;
            }
            default: Tb = TP; tVIC = 0.0; break;
        }
    }
    Tb = fmax(Tb, TMin);
    if (TMax < ICS_GREAT) Tb = fmin(Tb, TMax);
    double eb = th.Cv * Tb;
    if (!fixesT) Tb = eb / th.Cv;
    double psib = 1.0 / (th.R * Tb);
    f[Q_P * NX + s] = pb;
    f[Q_T * NX + s] = Tb;
    f[Q_PSI * NX + s] = psib;
    f[Q_RHO * NX + s] = psib * pb;
    f[Q_UX * NX + s] = Ub[0]; f[Q_UY * NX + s] = Ub[1]; f[Q_UZ * NX + s] = Ub[2];
    vic[b] = pVIC; vic[NB + b] = uVIC[0]; vic[2 * (size_t)NB + b] = uVIC[1]; vic[3 * (size_t)NB + b] = uVIC[2]; vic[4 * (size_t)NB + b] = tVIC;
    vic[5 * (size_t)NB + b] = gU[0]; vic[6 * (size_t)NB + b] = gU[1]; vic[7 * (size_t)NB + b] = gU[2]; vic[8 * (size_t)NB + b] = gT;
}

// derived fields E, H, cR, c on cell positions and physical boundary slots
__global__ void k_primitives(int NP, int NH, int NB, const int* __restrict__ pos2cell, const int* __restrict__ bfPatch,
                             const BCDev* __restrict__ bcs, Thermo th, int scheme, double* __restrict__ f, size_t NX)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    size_t s;
    if (i < NP) { if (pos2cell[i] < 0) return; s = i; }
    else {
        int b = i - NP;
        if (b >= NB) return;
        int pi = bfPatch[b];
        if (pi < 0 || bcs[pi].kind[0] == ICSB200_BC_COUPLED) return;
        s = (size_t)NP + NH + b;
    }
    double T = f[Q_T * NX + s], pr = f[Q_P * NX + s], rho = f[Q_RHO * NX + s], psi = f[Q_PSI * NX + s];
    double ux = f[Q_UX * NX + s], uy = f[Q_UY * NX + s], uz = f[Q_UZ * NX + s];
    double he = th.Cv * T;
    double E = he + 0.5 * (ux * ux + uy * uy + uz * uz);
    double H = fmax(E, ICS_SMALL) + fmax(pr / rho, ICS_SMALL);
    double cc = sqrt(th.gamma / psi);
    double cR;
    if (scheme == ICSB200_FLUX_AUSMPLUSUP) cR = sqrt(2.0 * (th.gamma - 1.0) / (th.gamma + 1.0) * H);
    else cR = fmax(cc, ICS_VSMALL);
    f[Q_E * NX + s] = E; f[Q_H * NX + s] = H; f[Q_C * NX + s] = cc; f[Q_CR * NX + s] = cR;
    // eCalc = rhoE/rho - 0.5 magSqr(U) (residualsUpdate.H:40); boundary rhoE as updateFields.H:99-104 builds it
    const double k2 = ux * ux + uy * uy + uz * uz;
    const double rhoE = (i < NP) ? f[Q_W4 * NX + s] : rho * (he + 0.5 * k2);
    const double rho0 = (i < NP) ? f[Q_W0 * NX + s] : rho;
    f[Q_EC * NX + s] = rhoE / rho0 - 0.5 * k2;
}

// updateFields.H:7-22 — W += dW, U, e; flag cells that need energy bounding (max(neg(e-eBound)) > 0.5, :46,:58)
__global__ void k_update1(int NP, const int* __restrict__ pos2cell, const double* __restrict__ dW, size_t NPH, double rhoMin, double eMin,
                          double eMax, double* __restrict__ f, size_t NX, double* __restrict__ eOut, int* __restrict__ flags)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    double rho = f[Q_W0 * NX + p] + dW[p];
    double rux = f[Q_W1 * NX + p] + dW[NPH + p], ruy = f[Q_W2 * NX + p] + dW[2 * NPH + p], ruz = f[Q_W3 * NX + p] + dW[3 * NPH + p];
    double rE = f[Q_W4 * NX + p] + dW[4 * NPH + p];
    if (rhoMin > -ICS_GREAT) rho = fmax(rho, rhoMin);
    double ux = rux / rho, uy = ruy / rho, uz = ruz / rho;
    double e = rE / rho - 0.5 * (ux * ux + uy * uy + uz * uz);
    f[Q_RHO * NX + p] = rho;
    f[Q_UX * NX + p] = ux; f[Q_UY * NX + p] = uy; f[Q_UZ * NX + p] = uz;
    eOut[p] = e;
    if (e - eMin < 0) atomicOr(&flags[0], 1);
    if (e - eMax >= 0) atomicOr(&flags[1], 1);
}

// updateFields.H:46-78 — bound e, thermo.correct() (T, psi), p = rho/psi, recompute rhoU, rhoE
__global__ void k_update2(int NP, const int* __restrict__ pos2cell, Thermo th, double eMin, double eMax, bool haveTMax,
                          const double* __restrict__ eIn, const int* __restrict__ flags, double* __restrict__ f, size_t NX)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    double e = eIn[p];
    if (flags[0]) e = fmax(e, eMin);
    if (haveTMax && flags[1]) e = fmin(e, eMax);
    double rho = f[Q_RHO * NX + p];
    double ux = f[Q_UX * NX + p], uy = f[Q_UY * NX + p], uz = f[Q_UZ * NX + p];
    double T = e / th.Cv;
    double psi = 1.0 / (th.R * T);
    f[Q_T * NX + p] = T;
    f[Q_PSI * NX + p] = psi;
    f[Q_P * NX + p] = rho / psi;
    f[Q_W0 * NX + p] = rho;
    f[Q_W1 * NX + p] = rho * ux; f[Q_W2 * NX + p] = rho * uy; f[Q_W3 * NX + p] = rho * uz;
    f[Q_W4 * NX + p] = rho * (e + 0.5 * (ux * ux + uy * uy + uz * uz));
}

// boundLocalTimeStep.H:10-45 — cells whose rho / e dropped below lowerBound x previous value
__global__ void k_bad(int NP, const int* __restrict__ pos2cell, double lb, const double* __restrict__ Wprev, size_t NPH, const double* __restrict__ f,
                      size_t NX, double* __restrict__ bad)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    if (pos2cell[p] < 0) { bad[p] = 0.0; return; }
    double rp = Wprev[p], rup[3] = {Wprev[NPH + p] / rp, Wprev[2 * NPH + p] / rp, Wprev[3 * NPH + p] / rp};
    double rhoMin = lb * rp;
    double eMin = lb * (Wprev[4 * NPH + p] / rp - 0.5 * (rup[0] * rup[0] + rup[1] * rup[1] + rup[2] * rup[2]));
    double r = f[Q_W0 * NX + p];
    double u[3] = {f[Q_W1 * NX + p] / r, f[Q_W2 * NX + p] / r, f[Q_W3 * NX + p] / r};
    double eTemp = f[Q_W4 * NX + p] / r - 0.5 * (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    bad[p] = ((r < rhoMin) || (eTemp < eMin) || (eTemp < ICS_SMALL)) ? 1.0 : 0.0;
}

// boundLocalTimeStep.H:47-97 — factor 0.5 for a bad cell, 0.75 next to one; pseudoCoField *= factor
__global__ void k_factor(int NP, const int* __restrict__ pos2cell, const int* __restrict__ sliceOff, const int* __restrict__ rowNAll,
                         const int* __restrict__ col, const int* __restrict__ meta, const double* __restrict__ bad, double* __restrict__ co)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)sliceOff[p >> 5];
    double factor = 1.0;
    bool hasFace = false, nbrBad = false;
    for (int j = 0; j < rowNAll[p]; j++) {
        size_t e = (base + j) * 32 + lane;
        if ((meta[e] & 3) == ET_PHYS) continue;
        hasFace = true;
        nbrBad |= bad[col[e]] != 0.0;
    }
    if (bad[p] != 0.0 && hasFace) factor = fmin(0.5, factor);
    if (nbrBad) factor = fmin(0.75, factor);
    co[p] *= factor;
}

__global__ void k_copy5(int NP, const double* __restrict__ src, size_t sstride, double* __restrict__ dst, size_t dstride)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    for (int k = 0; k < 5; k++) dst[k * dstride + p] = src[k * sstride + p];
}

__global__ void k_fill(int n, double v, double* __restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = v;
}

__global__ void k_gather_boundary(int NB, int off, const int* __restrict__ nbrSlot, const double* __restrict__ f, size_t NX, double* __restrict__ out)
{
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    // coupled faces report the patchNeighbourField (neighbour cell / halo slot), as the boundaryField of a coupled patch does
    size_t s = (nbrSlot && nbrSlot[b] >= 0) ? (size_t)nbrSlot[b] : (size_t)off + b;
    out[b] = f[Q_RHO * NX + s];
    out[NB + 3 * (size_t)b] = f[Q_UX * NX + s]; out[NB + 3 * (size_t)b + 1] = f[Q_UY * NX + s]; out[NB + 3 * (size_t)b + 2] = f[Q_UZ * NX + s];
    out[4 * (size_t)NB + b] = f[Q_P * NX + s];
    out[5 * (size_t)NB + b] = f[Q_T * NX + s];
}

}  // namespace

static Thermo thermoOf(const icsb200_ctx* c) { return Thermo{c->R, c->Cv, c->gamma}; }

int ics_eval_bc(icsb200_ctx* c, bool /*init*/)
{
    if (c->NB == 0) return 0;
    LaunchScope ls(c, TM_BC);
    k_bc<<<gridFor(c->NB, 128), 128, 0, c->stream>>>(c->NB, c->NP + c->NH, c->d_bfOwnerPos, c->d_bfPatch, c->d_bc, c->d_bfGeo, c->d_phiB,
                                                     thermoOf(c), c->sch.T_min, c->sch.T_max, c->d_fields, c->NX, c->d_vic);
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_primitives(icsb200_ctx* c)
{
    c->reconValid = false;  // the state changed: stored face reconstructions are stale
    {
        LaunchScope ls(c, TM_PRIM);
        int n = c->NP + c->NB;
        k_primitives<<<gridFor(n, 256), 256, 0, c->stream>>>(c->NP, c->NH, c->NB, c->d_pos2cell, c->d_bfPatch, c->d_bc, thermoOf(c),
                                                             c->sch.flux_scheme, c->d_fields, c->NX);
    }
    CUDA_TRY(c, cudaGetLastError());
    // neighbour-rank copies of the arrays the face kernels gather: rho p Ux Uy Uz cR E H c (+ eCalc) (contiguous ids 0..9)
    // + T (id 10) for the stored coupled-patch values of the full viscous Jacobian
    // phase-lag patches mix the time instances for rho p U cR E H c (ids 0..8), not for eCalc / T
    return ics_halo_fields(c, c->d_fields, c->NX, c->mu > 0 ? (c->sch.viscous_full_jacobian ? 11 : 10) : 9, 1u << Q_UX, 0x1FFu);
}

// conserved variables + boundary + derived fields from freshly uploaded p, U, T (host-facing iterate)
int ics_state_from_primitives(icsb200_ctx* c)
{
    {
        LaunchScope ls(c, TM_PRIM);
        k_state_init<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, thermoOf(c), c->d_fields, c->NX);
    }
    CUDA_TRY(c, cudaGetLastError());
    int r = ics_eval_bc(c, false);
    if (r) return r;
    return ics_primitives(c);
}

int ics_copy_prev(icsb200_ctx* c)
{
    LaunchScope ls(c, TM_VEC);
    k_copy5<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->q(Q_W0), c->NX, c->d_Wprev, c->NPH);
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_bound_local_dt(icsb200_ctx* c)
{
    if (!(c->sch.local_timestepping && c->sch.local_timestepping_bounding)) return 0;
    double* bad = c->d_w;  // scratch: one double per slot so the flags can ride the generic halo exchange
    {
        LaunchScope ls(c, TM_UPDATE);
        k_bad<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, c->sch.local_timestepping_lower_bound, c->d_Wprev, c->NPH,
                                                          c->d_fields, c->NX, bad);
    }
    CUDA_TRY(c, cudaGetLastError());
    // patchNeighbourField of the rho / e tests on processor patches (boundLocalTimeStep.H:60-95)
    int r = ics_halo_fields(c, bad, c->NPH, 1);
    if (r) return r;
    {
        LaunchScope ls(c, TM_UPDATE);
        k_factor<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_meta, bad, c->d_co);
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_update(icsb200_ctx* c)
{
    const double eMin = c->Cv * c->sch.T_min, eMax = c->Cv * c->sch.T_max;
    const bool haveTMax = c->sch.T_max < ICS_GREAT;
    int* flags = (int*)c->d_counter + 32;
    CUDA_TRY(c, cudaMemsetAsync(flags, 0, 2 * sizeof(int), c->stream));
    {
        LaunchScope ls(c, TM_UPDATE);
        k_update1<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_dW, c->NPH, c->sch.rho_min, eMin,
                                                              haveTMax ? eMax : ICS_VGREAT, c->d_fields, c->NX, c->d_w, flags);
    }
    // max(neg(e - eBound)) is a global reduction in the reference (updateFields.H:46,58)
    if (c->nRanks > 1 && ics_allreduce_max_int(c, flags, 2)) return ICSB200_ECUDA;
    {
        LaunchScope ls(c, TM_UPDATE);
        k_update2<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, thermoOf(c), eMin, eMax, haveTMax, c->d_w, flags,
                                                              c->d_fields, c->NX);
    }
    CUDA_TRY(c, cudaGetLastError());
    int r = ics_eval_bc(c, false);
    if (r) return r;
    return ics_primitives(c);
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int icsb200_state_set(icsb200_ctx* c, const double* p, const double* U, const double* T)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "state_set: mesh not set");
    if (!c->thermoSet) return ics_fail(c, ICSB200_ESTATE, "state_set: thermo not set");
    cudaSetDevice(c->device);
    int r;
    if ((r = ics_upload_cells(c, p, 1, c->q(Q_P), c->NX))) return r;
    if ((r = ics_upload_cells(c, U, 3, c->q(Q_UX), c->NX))) return r;
    if ((r = ics_upload_cells(c, T, 1, c->q(Q_T), c->NX))) return r;
    {
        LaunchScope ls(c, TM_PRIM);
        k_state_init<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, thermoOf(c), c->d_fields, c->NX);
    }
    if (c->NB > 0) {
        {
            LaunchScope ls(c, TM_BC);
            k_boundary_provisional<<<gridFor(c->NB, 128), 128, 0, c->stream>>>(c->NB, c->NP + c->NH, c->d_bfOwnerPos, c->d_bfPatch, c->d_bfGeo,
                                                                               c->d_fields, c->NX, c->d_phiB);
        }
        if ((r = ics_eval_bc(c, true))) return r;
        if ((r = ics_eval_bc(c, true))) return r;
        {
            LaunchScope ls(c, TM_BC);
            k_boundary_phi<<<gridFor(c->NB, 128), 128, 0, c->stream>>>(c->NB, c->NP + c->NH, c->d_bfPatch, c->d_bc, c->d_bfGeo, c->d_fields,
                                                                       c->NX, c->d_phiB);
        }
    }
    CUDA_TRY(c, cudaGetLastError());
    if ((r = ics_primitives(c))) return r;
    // old-time levels, pseudo time
    {
        LaunchScope ls(c, TM_VEC);
        k_copy5<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->q(Q_W0), c->NX, c->d_Wold, c->NP);
        k_copy5<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->q(Q_W0), c->NX, c->d_Wold2, c->NP);
        k_fill<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->sch.pseudo_co_num, c->d_co);
        k_fill<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, 0.0, c->d_rdt);
        c->launches += 3;
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->pseudoCoNum = c->sch.pseudo_co_num;
    c->timeIndex = 0;
    c->haveInitRes = c->havePrevRes = false;
    c->firstIter = true;
    c->stateSet = true;
    c->fluxValid = c->matrixSet = false;
    return 0;
}

extern "C" int icsb200_state_get(icsb200_ctx* c, double* rho, double* rhoU, double* rhoE, double* p, double* U, double* T)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "state_get: state not set");
    cudaSetDevice(c->device);
    int r = 0;
    if (rho && (r = ics_download_cells(c, rho, 1, c->q(Q_W0), c->NX))) return r;
    if (rhoU && (r = ics_download_cells(c, rhoU, 3, c->q(Q_W1), c->NX))) return r;
    if (rhoE && (r = ics_download_cells(c, rhoE, 1, c->q(Q_W4), c->NX))) return r;
    if (p && (r = ics_download_cells(c, p, 1, c->q(Q_P), c->NX))) return r;
    if (U && (r = ics_download_cells(c, U, 3, c->q(Q_UX), c->NX))) return r;
    if (T && (r = ics_download_cells(c, T, 1, c->q(Q_T), c->NX))) return r;
    return 0;
}

extern "C" int icsb200_boundary_get(icsb200_ctx* c, double* rho_b, double* U_b, double* p_b, double* T_b)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "boundary_get: state not set");
    const int NB = c->NB;
    if (NB == 0) return 0;
    int r = ics_ensure_stage(c, sizeof(double) * 6 * (size_t)NB);
    if (r) return r;
    int* d_nbr = nullptr;
    if (c->NH > 0 && (r = ics_halo_fields(c, c->q(Q_T), c->NX, 1))) return r;   // T is not part of the per-iteration halo set
    if ((int)c->bfNbrSlot.size() == NB) {
        CUDA_TRY(c, cudaMalloc((void**)&d_nbr, sizeof(int) * NB));
        CUDA_TRY(c, cudaMemcpyAsync(d_nbr, c->bfNbrSlot.data(), sizeof(int) * NB, cudaMemcpyHostToDevice, c->stream));
    }
    {
        LaunchScope ls(c, TM_PERM);
        k_gather_boundary<<<gridFor(NB, 128), 128, 0, c->stream>>>(NB, c->NP + c->NH, d_nbr, c->d_fields, c->NX, c->d_stage);
    }
    std::vector<double> h(6 * (size_t)NB);
    CUDA_TRY(c, cudaMemcpyAsync(h.data(), c->d_stage, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (d_nbr) cudaFree(d_nbr);
    if (rho_b) std::copy(h.begin(), h.begin() + NB, rho_b);
    if (U_b) std::copy(h.begin() + NB, h.begin() + 4 * (size_t)NB, U_b);
    if (p_b) std::copy(h.begin() + 4 * (size_t)NB, h.begin() + 5 * (size_t)NB, p_b);
    if (T_b) std::copy(h.begin() + 5 * (size_t)NB, h.end(), T_b);
    return 0;
}

extern "C" int icsb200_new_time_step(icsb200_ctx* c)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "new_time_step: state not set");
    {
        LaunchScope ls(c, TM_VEC);
        k_copy5<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_Wold, c->NP, c->d_Wold2, c->NP);
        k_copy5<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->q(Q_W0), c->NX, c->d_Wold, c->NP);
        c->launches += 1;
    }
    CUDA_TRY(c, cudaGetLastError());
    c->timeIndex++;
    c->haveInitRes = c->havePrevRes = false;  // beginTimeStep.H:1-6
    c->firstIter = true;
    return 0;
}

extern "C" int icsb200_update_fields(icsb200_ctx* c)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "update_fields: state not set");
    int r = ics_bound_local_dt(c);
    if (r) return r;
    r = ics_update(c);
    if (r) return r;
    c->firstIter = false;
    c->fluxValid = c->matrixSet = false;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}
