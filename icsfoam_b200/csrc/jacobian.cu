// jacobian.cu — pseudo time step and approximate flux Jacobian, assembled straight into the sliced-ELL block rows.
//
// Replaces convectiveFluxScheme::createConvectiveJacobian (convectiveFluxScheme.C:537-546):
//   addFluxTerms :374-484 (7 x blockFvMatrix::insertBlock, blockFvMatrix.C:211-268),
//   addDissipationJacobian :487-534 (3 x insertDissipationBlock, blockFvMatrix.C:271-326),
//   addBoundaryTerms + boundaryJacobian :47-120, 219-355, addTemporalTerms :358-369,
// the Lax-Friedrichs branch of viscousFluxScheme::addFluxTerms (viscousFluxScheme.C:220-246 with
// fvj::laplacian(sf, one), blockFvOperatorsTemplates.C:499-557), setCoAndDeltaT.H:39-173 (local pseudo time step)
// and outerLoop.H:61-64 (ddtCoeff).
//
// One thread per cell row.  The reference builds 10 temporary LDU matrices with face loops and scatters
// (negSumDiag); here the row owner gathers its faces in ascending face id, which visits the same addends in the
// same order, so every diag / upper / lower coefficient carries the reference's rounding.  Each face's two 5x5
// Jacobians (left state, right state) are evaluated by both adjacent rows: one lands in the row's off-diagonal
// block, the other in its diagonal — no atomics, no colouring, all stores coalesced across the slice.
#include "common.cuh"

namespace {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double magSqr(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
__device__ __forceinline__ V3 cm(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ double sgn(double s) { return s >= 0 ? 1.0 : -1.0; }

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

struct JacArgs {
    int NP, NB, F;
    const int *pos2cell, *sliceOff, *rowNAll, *col, *meta, *gfid, *bfKind;
    const double *geo, *dCoupled, *C, *V, *f, *grad, *vic, *co;
    size_t NFG, NX, NPH;
    int limU, limT, localDt, useStoredRdt;
    double gamma, Cv, rdtUniform;
    int ddtScheme;
    double rDeltaT, coefft;
    double mu, alphaEff;  // viscous LF Jacobian (mu > 0)
    double *offd, *diag, *rD, *rdt, *ddtCoeff;
    const double* recon;  // [8*NFG] limited face states stored by k_flux_faces (REUSE instantiation)
    const double* tr;              // [2][NX] muEff, alphaEff fields or null (laminar constants)
    double R, TMin;                // thermo / bounds for the values stored on coupled patches (full viscous Jacobian)
    const double *mrf, *mrfOmega;  // MRFFaceVelocity [NFG] (GPU face order) and MRFOmega [3*NP]; null = zero field
};

// analytic Euler flux Jacobian d(F.n)/dW at state (U, E) — convectiveFluxScheme.C:402-464
// J[r*5+c], variable order (rho, rhoU, rhoE)
__device__ __forceinline__ void eulerJacobian(V3 U, double E, V3 n, double gamma, double* J)
{
    const double theta = 0.5 * (gamma - 1) * magSqr(U);
    const double a1 = gamma * E - theta;
    const double a2 = gamma - 1;
    const double projU = dot(U, n);
    J[0] = 0.0; J[1] = n.x; J[2] = n.y; J[3] = n.z; J[4] = 0.0;
    const V3 mr = n * theta - U * projU;
    J[5] = mr.x; J[10] = mr.y; J[15] = mr.z;
    const V3 a2n = a2 * n;
    // U*n - a2*n*U + projU*I   (outer products)
    J[6] = (U.x * n.x - a2n.x * U.x) + projU * 1.0; J[7] = (U.x * n.y - a2n.x * U.y) + projU * 0.0; J[8] = (U.x * n.z - a2n.x * U.z) + projU * 0.0;
    J[11] = (U.y * n.x - a2n.y * U.x) + projU * 0.0; J[12] = (U.y * n.y - a2n.y * U.y) + projU * 1.0; J[13] = (U.y * n.z - a2n.y * U.z) + projU * 0.0;
    J[16] = (U.z * n.x - a2n.z * U.x) + projU * 0.0; J[17] = (U.z * n.y - a2n.z * U.y) + projU * 0.0; J[18] = (U.z * n.z - a2n.z * U.z) + projU * 1.0;
    const V3 me = n * a2;
    J[9] = me.x; J[14] = me.y; J[19] = me.z;
    J[20] = projU * (theta - a1);
    const V3 er = n * a1 - a2 * U * projU;
    J[21] = er.x; J[22] = er.y; J[23] = er.z;
    J[24] = gamma * projU;
}

__device__ __forceinline__ bool isDiagEntry(int k) { return k == 0 || k == 6 || k == 12 || k == 18 || k == 24; }

// FULLV: LaxFriedrichJacobian false — viscousFluxScheme::addFluxTerms :248-261 (five fvj::laplacian(sf, vf) blocks,
// blockFvOperatorsTemplates.C:387-450,575-637: upper = vf[nei] sf2, lower = vf[own] sf2, negSumDiag, coupled: diag[own] -=
// vf[own] sf2_b, interfacesUpper = vf_b sf2_b with the values STORED on the patch) and addBoundaryTerms :120-215
template <bool REUSE, bool FULLV>
__global__ void __launch_bounds__(128, 3)
k_jac(JacArgs a)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.NP || a.pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)a.sliceOff[p >> 5];
    const int nAll = a.rowNAll[p];
    const bool lim4U = a.limU == ICSB200_LIM_VANLEER || a.limU == ICSB200_LIM_MINMOD;
    const bool lim4T = a.limT == ICSB200_LIM_VANLEER || a.limT == ICSB200_LIM_MINMOD;
    double conv[25], dissAcc = 0.0, viscAcc = 0.0, mrfAcc = 0.0, rdt = 0.0;
#pragma unroll
    for (int k = 0; k < 25; k++) conv[k] = 0.0;
    bool hasPhys = false;
    const double cOwn = a.f[(size_t)Q_C * a.NX + p];
    const V3 UOwn = {a.f[(size_t)Q_UX * a.NX + p], a.f[(size_t)Q_UY * a.NX + p], a.f[(size_t)Q_UZ * a.NX + p]};
    // full viscous Jacobian: vf = -U/rho, 1/rho, -E/rho + |U|^2/rho at the row cell and the negSumDiag accumulators of the five blocks
    double vRowU[3] = {0, 0, 0}, vRowInv = 0.0, vRowE = 0.0, aMomRho[3] = {0, 0, 0}, aMomU = 0.0, aERho = 0.0, aEU[3] = {0, 0, 0}, aEE = 0.0;
    if (FULLV) {
        const double rho = a.f[(size_t)Q_RHO * a.NX + p], E = a.f[(size_t)Q_E * a.NX + p];
        vRowU[0] = -UOwn.x / rho; vRowU[1] = -UOwn.y / rho; vRowU[2] = -UOwn.z / rho;
        vRowInv = 1.0 / rho;
        vRowE = -E / rho + (UOwn.x * UOwn.x + UOwn.y * UOwn.y + UOwn.z * UOwn.z) / rho;
    }
    for (int j = 0; j < nAll; j++) {
        const size_t e = (base + j) * 32 + lane;
        const int c = a.col[e], m = a.meta[e], type = m & 3;
        const size_t g = a.gfid[e];
        const int b = (m >> 2) - a.F;
        const V3 Sf = {a.geo[G_SFX * a.NFG + g], a.geo[G_SFY * a.NFG + g], a.geo[G_SFZ * a.NFG + g]};
        const double magSf = a.geo[G_MAGSF * a.NFG + g];
        const V3 n = Sf / magSf;
        double* blk = a.offd + ((base + j) * 25) * 32 + lane;
        const double mrf = a.mrf ? a.mrf[g] : 0.0;  // flux.MRFFaceVelocity() of this face
        if (type == ET_PHYS) {
            hasPhys = true;
            // lambdaConv boundary value: c_b + |U_b & n|  (interpolate() returns the patch value)
            const V3 Ub = {a.f[(size_t)Q_UX * a.NX + c], a.f[(size_t)Q_UY * a.NX + c], a.f[(size_t)Q_UZ * a.NX + c]};
            const double lam = a.f[(size_t)Q_C * a.NX + c] + fabs(dot(Ub, n) - mrf);
            const double dl = 0.5 * magSf * lam;
            dissAcc -= dl;  // mx.diag[own] -= interfacesLower (physical boundaries included, blockFvMatrix.C:310-321)
            if (a.bfKind[b] == ICSB200_WALL) {
                // setCoAndDeltaT.H:87-138: wall patches use the cell state with a factor 1/2
                const double pLambda = 0.5 * a.geo[G_NONORTH * a.NFG + g] * (cOwn + fabs(dot(UOwn, n) - mrf));
                rdt = fmax(rdt, pLambda);
            }
#pragma unroll
            for (int k = 0; k < 25; k++) blk[(size_t)k * 32] = 0.0;
            continue;
        }
        const bool rowIsP = type != ET_LOWER;
        const bool coupled = type == ET_COUPLED;
        const int P = rowIsP ? p : c, N = rowIsP ? c : p;
        const double w = a.geo[G_W * a.NFG + g];
        V3 d;
        if (coupled) d = {a.dCoupled[b], a.dCoupled[a.NB + b], a.dCoupled[2 * (size_t)a.NB + b]};
        else d = {a.C[N] - a.C[P], a.C[a.NPH + N] - a.C[a.NPH + P], a.C[2 * a.NPH + N] - a.C[2 * a.NPH + P]};
        // limited reconstruction of U (3 comps) and E on both sides (convectiveFluxScheme.C:387-400)
        double L[4], R[4];
        if (REUSE) {
            // the limited L/R states of this face were stored by the flux kernel for this very state
            L[0] = a.recon[g]; L[1] = a.recon[a.NFG + g]; L[2] = a.recon[2 * a.NFG + g];
            R[0] = a.recon[3 * a.NFG + g]; R[1] = a.recon[4 * a.NFG + g]; R[2] = a.recon[5 * a.NFG + g];
            L[3] = a.recon[6 * a.NFG + g]; R[3] = a.recon[7 * a.NFG + g];
        } else
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = (q < 3) ? (Q_UX + q) : Q_E;
            const int lim = (q < 3) ? a.limU : a.limT;
            const bool useG = (q < 3) ? lim4U : lim4T;
            const double phiP = a.f[(size_t)k * a.NX + P], phiN = a.f[(size_t)k * a.NX + N];
            double gcP = 0.0, gcN = 0.0;
            if (useG) {
                gcP = d.x * a.grad[(size_t)(k * 3) * a.NPH + P] + d.y * a.grad[(size_t)(k * 3 + 1) * a.NPH + P] + d.z * a.grad[(size_t)(k * 3 + 2) * a.NPH + P];
                gcN = d.x * a.grad[(size_t)(k * 3) * a.NPH + N] + d.y * a.grad[(size_t)(k * 3 + 1) * a.NPH + N] + d.z * a.grad[(size_t)(k * 3 + 2) * a.NPH + N];
            }
            const double gradf = phiN - phiP;
            const double limL = limiterOf(lim, gcP, gradf), limR = limiterOf(lim, gcN, gradf);
            const double wL = limL * w + (1.0 - limL) * 1.0, wR = limR * w + (1.0 - limR) * 0.0;
            if (coupled) { L[q] = wL * phiP + (1.0 - wL) * phiN; R[q] = wR * phiP + (1.0 - wR) * phiN; }
            else { L[q] = wL * (phiP - phiN) + phiN; R[q] = wR * (phiP - phiN) + phiN; }
        }
        // lambdaConv = interpolate(c) + |interpolate(U) & n - MRFFaceVelocity|   (convectiveFluxScheme.C:498-524)
        double lam;
        {
            const double cP = a.f[(size_t)Q_C * a.NX + P], cN = a.f[(size_t)Q_C * a.NX + N];
            const V3 UP = {a.f[(size_t)Q_UX * a.NX + P], a.f[(size_t)Q_UY * a.NX + P], a.f[(size_t)Q_UZ * a.NX + P]};
            const V3 UN = {a.f[(size_t)Q_UX * a.NX + N], a.f[(size_t)Q_UY * a.NX + N], a.f[(size_t)Q_UZ * a.NX + N]};
            double cf;
            V3 uf;
            if (coupled) {
                cf = w * cP + (1.0 - w) * cN;
                uf = {w * UP.x + (1.0 - w) * UN.x, w * UP.y + (1.0 - w) * UN.y, w * UP.z + (1.0 - w) * UN.z};
            } else {
                cf = w * (cP - cN) + cN;
                uf = {w * (UP.x - UN.x) + UN.x, w * (UP.y - UN.y) + UN.y, w * (UP.z - UN.z) + UN.z};
            }
            lam = cf + fabs(dot(uf, n) - mrf);
        }
        // MRFdivMeshPhi = fvj::div(w, MRFFaceVelocity*magSf) (convectiveFluxScheme.C:477-481, blockFvOperatorsTemplates.C:146-201):
        // upper = sf (1 - w), lower = -sf w, negSumDiag; subtracted from the three diagonal-variable blocks
        double mrfOff = 0.0;
        if (a.mrf) {
            const double sfm = mrf * magSf;
            const double mU = sfm * (1 - w), mL = -sfm * w;
            if (rowIsP) { mrfOff = mU; mrfAcc -= mL; }
            else { mrfOff = mL; mrfAcc -= mU; }
        }
        rdt = fmax(rdt, a.geo[G_NONORTH * a.NFG + g] * lam);
        const double dl = 0.5 * magSf * lam;
        dissAcc -= dl;
        double sf2 = 0.0, sf2mu = 0.0, sf2al = 0.0, vColU[3] = {0, 0, 0}, vColInv = 0.0, vColE = 0.0;
        if (a.mu > 0) {
            const double rP = a.f[(size_t)Q_RHO * a.NX + P], rN = a.f[(size_t)Q_RHO * a.NX + N];
            const double rhof = coupled ? (w * rP + (1.0 - w) * rN) : (w * (rP - rN) + rN);
            // fvc::interpolate(turbulence.muEff()), fvc::interpolate(turbulence.alphaEff()) (viscousFluxScheme.C:222-223)
            const double muP = a.tr ? a.tr[P] : a.mu, muN = a.tr ? a.tr[N] : a.mu;
            const double alP = a.tr ? a.tr[a.NX + P] : a.alphaEff, alN = a.tr ? a.tr[a.NX + N] : a.alphaEff;
            const double muf = coupled ? (w * muP + (1.0 - w) * muN) : (w * (muP - muN) + muN);
            const double alf = coupled ? (w * alP + (1.0 - w) * alN) : (w * (alP - alN) + alN);
            if (!FULLV) {
                const double lambdaVisc = (muf + alf) / rhof;
                sf2 = (0.5 * lambdaVisc) * magSf * a.geo[G_DELTA * a.NFG + g];
                viscAcc -= sf2;
            } else {
                sf2mu = muf * magSf * a.geo[G_DELTA * a.NFG + g];
                sf2al = alf * magSf * a.geo[G_DELTA * a.NFG + g];
                // vf at the column: the neighbour cell, or the values stored on the coupled patch (updateFields.H:80-104):
                // cyclic / cyclicAMI store w*internal + (1-w)*patchNeighbourField, processor patches the neighbour values;
                // then T = THE(he(T)), psi = 1/(R T), rho = psi p, E = he + 0.5 |U|^2
                double rhoC, EC, UC[3] = {a.f[(size_t)Q_UX * a.NX + c], a.f[(size_t)Q_UY * a.NX + c], a.f[(size_t)Q_UZ * a.NX + c]};
                if (!coupled) {
                    rhoC = a.f[(size_t)Q_RHO * a.NX + c];
                    EC = a.f[(size_t)Q_E * a.NX + c];
                } else {
                    double pb = a.f[(size_t)Q_P * a.NX + c], Tb = a.f[(size_t)Q_T * a.NX + c];
                    if (a.bfKind[b] != ICSB200_PROCESSOR) {
                        pb = w * a.f[(size_t)Q_P * a.NX + p] + (1.0 - w) * pb;
                        Tb = w * a.f[(size_t)Q_T * a.NX + p] + (1.0 - w) * Tb;
                        UC[0] = w * UOwn.x + (1.0 - w) * UC[0]; UC[1] = w * UOwn.y + (1.0 - w) * UC[1]; UC[2] = w * UOwn.z + (1.0 - w) * UC[2];
                    }
                    Tb = fmax(Tb, a.TMin);
                    const double eb = a.Cv * Tb;
                    Tb = eb / a.Cv;
                    const double psib = 1.0 / (a.R * Tb);
                    rhoC = psib * pb;
                    EC = a.Cv * Tb + 0.5 * (UC[0] * UC[0] + UC[1] * UC[1] + UC[2] * UC[2]);
                }
                vColU[0] = -UC[0] / rhoC; vColU[1] = -UC[1] / rhoC; vColU[2] = -UC[2] / rhoC;
                vColInv = 1.0 / rhoC;
                vColE = -EC / rhoC + (UC[0] * UC[0] + UC[1] * UC[1] + UC[2] * UC[2]) / rhoC;
#pragma unroll
                for (int d = 0; d < 3; d++) { aMomRho[d] -= vRowU[d] * sf2mu; aEU[d] -= vRowU[d] * sf2al; }
                aMomU -= vRowInv * sf2mu; aERho -= vRowE * sf2al; aEE -= vRowInv * sf2al;
            }
        }
        double JL[25], JR[25];
        eulerJacobian({L[0], L[1], L[2]}, L[3], n, a.gamma, JL);
        eulerJacobian({R[0], R[1], R[2]}, R[3], n, a.gamma, JR);
        const double hp = 0.5 * magSf, hm = -0.5 * magSf;
#pragma unroll
        for (int k = 0; k < 25; k++) {
            const bool hasConv = !(k == 0 || k == 4);  // dSByS(0,0) and dSByS(0,1) receive no insertBlock
            const double upp = hp * JR[k];             // upper / interfacesUpper
            const double low = hm * JL[k];             // lower / interfacesLower
            double off;
            if (rowIsP) { off = hasConv ? (0.0 + upp) : 0.0; if (hasConv) conv[k] -= low; }
            else { off = hasConv ? (0.0 + low) : 0.0; if (hasConv) conv[k] -= upp; }
            if (isDiagEntry(k)) { if (a.mrf) off -= mrfOff * 1.0; off -= dl; if (!FULLV && a.mu > 0) off -= sf2 * 1.0; }
            if (FULLV) {
                if (k == 5 || k == 10 || k == 15) off -= vColU[k / 5 - 1] * sf2mu;        // dMomByRho
                else if (k == 6 || k == 12 || k == 18) off -= vColInv * sf2mu;            // dMomByRhoU (* I)
                else if (k == 20) off -= vColE * sf2al;                                   // dEnergyByRho
                else if (k >= 21 && k <= 23) off -= vColU[k - 21] * sf2al;                // dEnergyByRhoU
                else if (k == 24) off -= vColInv * sf2al;                                 // dEnergyByRhoE
            }
            blk[(size_t)k * 32] = off;
        }
    }
    // ---- pseudo time step (setCoAndDeltaT.H:57-147) and ddtCoeff (outerLoop.H:61-64)
    const double vol = a.V[p];
    if (a.useStoredRdt) rdt = a.rdt[p];
    else if (a.localDt) rdt = rdt / a.co[p];
    else rdt = a.rdtUniform;
    double innerDiag = 0.0;
    if (a.ddtScheme == ICSB200_DDT_EULER) innerDiag = a.rDeltaT * vol;
    else if (a.ddtScheme == ICSB200_DDT_BACKWARD) innerDiag = (a.coefft * a.rDeltaT) * vol;
    const double ddtCoeff = (a.ddtScheme == ICSB200_DDT_STEADY ? (0.0 + rdt * vol) : (innerDiag + rdt * vol)) / vol;
    a.rdt[p] = rdt;
    a.ddtCoeff[p] = ddtCoeff;
    // ---- diagonal block: conv (0 + mx.diag), -= dissipation mx.diag, += boundary terms, += temporal, -= viscous
    double dg[25];
#pragma unroll
    for (int k = 0; k < 25; k++) {
        dg[k] = 0.0 + conv[k];
        if (isDiagEntry(k)) { if (a.mrf) dg[k] -= mrfAcc * 1.0; dg[k] -= dissAcc; }
    }
    if (hasPhys) {
        const double rhoI = a.f[(size_t)Q_RHO * a.NX + p];
        const double TI = a.f[(size_t)Q_T * a.NX + p];
        const V3 UI = UOwn;
        const double rhoEI = rhoI * (a.Cv * TI + 0.5 * magSqr(UI));
        const double gammaI = a.gamma, cvI = a.Cv;
        const double dPdRho = 0.5 * (gammaI - 1) * magSqr(UI);
        const V3 dUdRho = -1.0 * UI / rhoI;
        const double dTdRho = -1.0 / (cvI * rhoI) * (rhoEI / rhoI - magSqr(UI));
        const V3 dPdRhoU = -(gammaI - 1) * UI;
        const double dUdRhoU = 1.0 / rhoI * 1.0;
        const V3 dTdRhoU = -1.0 * UI / (cvI * rhoI);
        const double dPdRhoE = gammaI - 1;
        const double dTdRhoE = 1.0 / (cvI * rhoI);
        for (int j = 0; j < nAll; j++) {
            const size_t e = (base + j) * 32 + lane;
            const int m = a.meta[e];
            if ((m & 3) != ET_PHYS) continue;
            const int c = a.col[e];
            const size_t g = a.gfid[e];
            const int b = (m >> 2) - a.F;
            const V3 SfB = {a.geo[G_SFX * a.NFG + g], a.geo[G_SFY * a.NFG + g], a.geo[G_SFZ * a.NFG + g]};
            const double pVIC = a.vic[b], tVIC = a.vic[4 * (size_t)a.NB + b];
            const V3 uVIC = {a.vic[a.NB + b], a.vic[2 * (size_t)a.NB + b], a.vic[3 * (size_t)a.NB + b]};
            const double rhoB = a.f[(size_t)Q_RHO * a.NX + c], pB = a.f[(size_t)Q_P * a.NX + c], TB = a.f[(size_t)Q_T * a.NX + c];
            const V3 UB = {a.f[(size_t)Q_UX * a.NX + c], a.f[(size_t)Q_UY * a.NX + c], a.f[(size_t)Q_UZ * a.NX + c]};
            double UrelBdotSf = dot(UB, SfB);
            UrelBdotSf -= (a.mrf ? a.mrf[g] : 0.0) * a.geo[G_MAGSF * a.NFG + g];
            const double rhoEB = rhoB * (a.Cv * TB + 0.5 * magSqr(UB));
            const double cvB = a.Cv;
            // boundaryJacobian (convectiveFluxScheme.C:96-117)
            const double dContFluxdp = rhoB / pB * UrelBdotSf * pVIC;
            const V3 dContFluxdU = rhoB * cm(SfB, uVIC);
            const double dContFluxdT = -rhoB / TB * UrelBdotSf * tVIC;
            const V3 dMomFluxdp = (rhoB / pB * UB * UrelBdotSf + SfB) * pVIC;
            const V3 rUB = rhoB * UB, sv = cm(SfB, uVIC);
            double dMomFluxdU[9] = {rUB.x * sv.x, rUB.x * sv.y, rUB.x * sv.z, rUB.y * sv.x, rUB.y * sv.y, rUB.y * sv.z, rUB.z * sv.x, rUB.z * sv.y, rUB.z * sv.z};
            const V3 dMomFluxdUDiag = rhoB * UrelBdotSf * uVIC;
            const V3 dMomFluxdT = -rhoB / TB * UB * UrelBdotSf * tVIC;
            const double dEnergyFluxdp = (rhoEB / pB * UrelBdotSf + dot(UB, SfB)) * pVIC;
            const V3 dEnergyFluxdU = cm(SfB, uVIC) * (rhoEB + pB) + rhoB * UrelBdotSf * cm(UB, uVIC);
            const double dEnergyFluxdT = UrelBdotSf * (rhoB * cvB - rhoEB / TB) * tVIC;
            dMomFluxdU[0] = dMomFluxdU[0] + dMomFluxdUDiag.x;
            dMomFluxdU[4] = dMomFluxdU[4] + dMomFluxdUDiag.y;
            dMomFluxdU[8] = dMomFluxdU[8] + dMomFluxdUDiag.z;
            // addBoundaryTerms (convectiveFluxScheme.C:293-349): three separate += per sub-block
            dg[0] += dContFluxdp * dPdRho;
            dg[0] += dot(dContFluxdU, dUdRho);
            dg[0] += dContFluxdT * dTdRho;
            { const V3 t1 = dContFluxdp * dPdRhoU, t2 = dContFluxdU * dUdRhoU, t3 = dContFluxdT * dTdRhoU;
              dg[1] += t1.x; dg[2] += t1.y; dg[3] += t1.z; dg[1] += t2.x; dg[2] += t2.y; dg[3] += t2.z; dg[1] += t3.x; dg[2] += t3.y; dg[3] += t3.z; }
            dg[4] += dContFluxdp * dPdRhoE;
            dg[4] += dContFluxdU.x * 0.0 + dContFluxdU.y * 0.0 + dContFluxdU.z * 0.0;
            dg[4] += dContFluxdT * dTdRhoE;
            { const V3 t1 = dMomFluxdp * dPdRho;
              const V3 t2 = {dMomFluxdU[0] * dUdRho.x + dMomFluxdU[1] * dUdRho.y + dMomFluxdU[2] * dUdRho.z,
                             dMomFluxdU[3] * dUdRho.x + dMomFluxdU[4] * dUdRho.y + dMomFluxdU[5] * dUdRho.z,
                             dMomFluxdU[6] * dUdRho.x + dMomFluxdU[7] * dUdRho.y + dMomFluxdU[8] * dUdRho.z};
              const V3 t3 = dMomFluxdT * dTdRho;
              dg[5] += t1.x; dg[10] += t1.y; dg[15] += t1.z; dg[5] += t2.x; dg[10] += t2.y; dg[15] += t2.z; dg[5] += t3.x; dg[10] += t3.y; dg[15] += t3.z; }
            {
                const double mp[3] = {dMomFluxdp.x, dMomFluxdp.y, dMomFluxdp.z}, pu[3] = {dPdRhoU.x, dPdRhoU.y, dPdRhoU.z};
                const double mt[3] = {dMomFluxdT.x, dMomFluxdT.y, dMomFluxdT.z}, tu[3] = {dTdRhoU.x, dTdRhoU.y, dTdRhoU.z};
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int cc = 0; cc < 3; cc++) {
                        const int k = (r + 1) * 5 + (cc + 1);
                        dg[k] += mp[r] * pu[cc];
                        dg[k] += dMomFluxdU[r * 3 + cc] * dUdRhoU;
                        dg[k] += mt[r] * tu[cc];
                    }
            }
            { const V3 t1 = dMomFluxdp * dPdRhoE;
              const V3 t2 = {dMomFluxdU[0] * 0.0 + dMomFluxdU[1] * 0.0 + dMomFluxdU[2] * 0.0, dMomFluxdU[3] * 0.0 + dMomFluxdU[4] * 0.0 + dMomFluxdU[5] * 0.0,
                             dMomFluxdU[6] * 0.0 + dMomFluxdU[7] * 0.0 + dMomFluxdU[8] * 0.0};
              const V3 t3 = dMomFluxdT * dTdRhoE;
              dg[9] += t1.x; dg[14] += t1.y; dg[19] += t1.z; dg[9] += t2.x; dg[14] += t2.y; dg[19] += t2.z; dg[9] += t3.x; dg[14] += t3.y; dg[19] += t3.z; }
            dg[20] += dEnergyFluxdp * dPdRho;
            dg[20] += dot(dEnergyFluxdU, dUdRho);
            dg[20] += dEnergyFluxdT * dTdRho;
            { const V3 t1 = dEnergyFluxdp * dPdRhoU, t2 = dEnergyFluxdU * dUdRhoU, t3 = dEnergyFluxdT * dTdRhoU;
              dg[21] += t1.x; dg[22] += t1.y; dg[23] += t1.z; dg[21] += t2.x; dg[22] += t2.y; dg[23] += t2.z; dg[21] += t3.x; dg[22] += t3.y; dg[23] += t3.z; }
            dg[24] += dEnergyFluxdp * dPdRhoE;
            dg[24] += dEnergyFluxdU.x * 0.0 + dEnergyFluxdU.y * 0.0 + dEnergyFluxdU.z * 0.0;
            dg[24] += dEnergyFluxdT * dTdRhoE;
        }
    }
    const double diagCoeff = ddtCoeff * vol;
    dg[0] += diagCoeff; dg[6] += diagCoeff * 1.0; dg[12] += diagCoeff * 1.0; dg[18] += diagCoeff * 1.0; dg[24] += diagCoeff;
    if (a.mrfOmega) {  // addMRFSource, diagonal part (convectiveFluxScheme.C:130-137)
        const double ox = a.mrfOmega[p], oy = a.mrfOmega[(size_t)a.NP + p], oz = a.mrfOmega[2 * (size_t)a.NP + p];
        dg[7] -= oz * vol; dg[8] += oy * vol; dg[11] += oz * vol; dg[13] -= ox * vol; dg[16] -= oy * vol; dg[17] += ox * vol;
    }
    if (!FULLV && a.mu > 0) { dg[0] -= viscAcc * 1.0; dg[6] -= viscAcc * 1.0; dg[12] -= viscAcc * 1.0; dg[18] -= viscAcc * 1.0; dg[24] -= viscAcc * 1.0; }
    if (FULLV) {
        dg[5] -= aMomRho[0]; dg[10] -= aMomRho[1]; dg[15] -= aMomRho[2];
        dg[6] -= aMomU; dg[12] -= aMomU; dg[18] -= aMomU;
        dg[20] -= aERho; dg[21] -= aEU[0]; dg[22] -= aEU[1]; dg[23] -= aEU[2]; dg[24] -= aEE;
        if (hasPhys) {
            // viscousFluxScheme::addBoundaryTerms (:120-215): wall terms through gradientInternalCoeffs of U and T
            const double rhoI = a.f[(size_t)Q_RHO * a.NX + p], TI = a.f[(size_t)Q_T * a.NX + p], cvI = a.Cv;
            const V3 UI = UOwn;
            const double EI = a.Cv * TI + 0.5 * magSqr(UI);
            const V3 dUdRho = -1.0 * UI / rhoI;
            const double dTdRho = -1.0 / (cvI * rhoI) * (EI - magSqr(UI));
            const double dUdRhoU = 1.0 / rhoI * 1.0;
            const V3 dTdRhoU = -1.0 * UI / (cvI * rhoI);
            const double dTdRhoE = 1.0 / (cvI * rhoI);
            const double du[3] = {dUdRho.x, dUdRho.y, dUdRho.z}, dtu[3] = {dTdRhoU.x, dTdRhoU.y, dTdRhoU.z};
            for (int j = 0; j < nAll; j++) {
                const size_t e = (base + j) * 32 + lane;
                const int m = a.meta[e];
                if ((m & 3) != ET_PHYS) continue;
                const int c = a.col[e];
                const size_t g = a.gfid[e];
                const int b = (m >> 2) - a.F;
                const double magSfB = a.geo[G_MAGSF * a.NFG + g], dcB = a.geo[G_DELTA * a.NFG + g];
                const double muB = a.tr ? a.tr[c] : a.mu, alB = a.tr ? a.tr[a.NX + c] : a.alphaEff;
                const double TGIC = -dcB * a.vic[8 * (size_t)a.NB + b];
                const double dEnergyFluxdT = -alB * a.Cv * TGIC * magSfB;
#pragma unroll
                for (int d = 0; d < 3; d++) {
                    const double uGIC = -dcB * a.vic[(size_t)(5 + d) * a.NB + b];
                    const double dMomFluxdU = -muB * uGIC * magSfB;
                    dg[5 + 5 * d] += dMomFluxdU * du[d];
                    dg[6 + 6 * d] += dMomFluxdU * dUdRhoU;
                    dg[9 + 5 * d] += dMomFluxdU * 0.0;
                }
                dg[20] += dEnergyFluxdT * dTdRho;
                dg[21] += dEnergyFluxdT * dtu[0]; dg[22] += dEnergyFluxdT * dtu[1]; dg[23] += dEnergyFluxdT * dtu[2];
                dg[24] += dEnergyFluxdT * dTdRhoE;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 25; k++) a.diag[(size_t)k * a.NP + p] = dg[k];
    // lusgs ctor (lusgs.C:50-125): rDiagCoeff = 1/max(|diag entries of the S-S and V-V diagonal blocks|)
    double rD = ICS_GREAT;
    rD = 1.0 / fmax(1.0 / rD, fabs(dg[0]));
    rD = 1.0 / fmax(1.0 / rD, fabs(dg[24]));
    rD = 1.0 / fmax(1.0 / rD, fabs(dg[6]));
    rD = 1.0 / fmax(1.0 / rD, fabs(dg[12]));
    rD = 1.0 / fmax(1.0 / rD, fabs(dg[18]));
    a.rD[p] = rD;
}

// non-local time stepping: max over faces of deltaCoeffs*lambda (setCoAndDeltaT.H:143-146)
__global__ void k_lambda_max(int NP, const int* __restrict__ pos2cell, const int* __restrict__ sliceOff, const int* __restrict__ rowNAll,
                             const int* __restrict__ rowNLow, const int* __restrict__ col, const int* __restrict__ meta, const int* __restrict__ gfid,
                             const double* __restrict__ geo, size_t NFG, const double* __restrict__ f, size_t NX, const double* __restrict__ mrf,
                             unsigned long long* __restrict__ out)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    double mx = 0.0;
    if (p < NP && pos2cell[p] >= 0) {
        const int lane = p & 31;
        const size_t base = (size_t)sliceOff[p >> 5];
        for (int j = rowNLow[p]; j < rowNAll[p]; j++) {
            const size_t e = (base + j) * 32 + lane;
            const int c = col[e], type = meta[e] & 3;
            const size_t g = gfid[e];
            const double magSf = geo[G_MAGSF * NFG + g], w = geo[G_W * NFG + g];
            const V3 n = V3{geo[G_SFX * NFG + g], geo[G_SFY * NFG + g], geo[G_SFZ * NFG + g]} / magSf;
            double cf;
            V3 uf;
            const double cP = f[(size_t)Q_C * NX + p], cN = f[(size_t)Q_C * NX + c];
            const V3 UP = {f[(size_t)Q_UX * NX + p], f[(size_t)Q_UY * NX + p], f[(size_t)Q_UZ * NX + p]};
            const V3 UN = {f[(size_t)Q_UX * NX + c], f[(size_t)Q_UY * NX + c], f[(size_t)Q_UZ * NX + c]};
            if (type == ET_PHYS) { cf = cN; uf = UN; }
            else if (type == ET_COUPLED) { cf = w * cP + (1.0 - w) * cN; uf = {w * UP.x + (1.0 - w) * UN.x, w * UP.y + (1.0 - w) * UN.y, w * UP.z + (1.0 - w) * UN.z}; }
            else { cf = w * (cP - cN) + cN; uf = {w * (UP.x - UN.x) + UN.x, w * (UP.y - UN.y) + UN.y, w * (UP.z - UN.z) + UN.z}; }
            mx = fmax(mx, geo[G_DELTA * NFG + g] * (cf + fabs(dot(uf, n) - (mrf ? mrf[g] : 0.0))));
        }
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(mx));
}

// ---- LDU views of the block rows (parity API) ----
__device__ __forceinline__ int blockRows(int blk, int* r0) { int rows; switch (blk) { case 0: case 1: case 4: *r0 = 0; rows = 1; break; case 2: case 3: case 5: *r0 = 4; rows = 1; break; default: *r0 = 1; rows = 3; } return rows; }
__device__ __forceinline__ int blockCols(int blk, int* c0) { int cols; switch (blk) { case 0: case 2: case 6: *c0 = 0; cols = 1; break; case 1: case 3: case 7: *c0 = 4; cols = 1; break; default: *c0 = 1; cols = 3; } return cols; }

// thread per (entry slot): copy sub-block coefficients of internal faces to upper[F*nc] / lower[F*nc] (or back)
__global__ void k_ldu_offdiag(long long nE32, const int* __restrict__ meta, const int* __restrict__ rowNAllBySlot, int blk, int F, bool toHost,
                              double* __restrict__ offd, double* __restrict__ upper, double* __restrict__ lower)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nE32) return;
    (void)rowNAllBySlot;
    const int m = meta[i], type = m & 3, fc = m >> 2;
    if (type > ET_UPPER || fc >= F || fc < 0) return;
    const long long ent = i >> 5;
    const int lane = (int)(i & 31);
    int r0, c0;
    const int rows = blockRows(blk, &r0), cols = blockCols(blk, &c0), nc = rows * cols;
    double* dst = (type == ET_UPPER) ? upper : lower;
    if (!dst) return;
    for (int r = 0; r < rows; r++)
        for (int cc = 0; cc < cols; cc++) {
            double* v = offd + ((size_t)ent * 25 + (r0 + r) * 5 + (c0 + cc)) * 32 + lane;
            if (toHost) dst[(size_t)fc * nc + r * cols + cc] = *v;
            else *v = dst[(size_t)fc * nc + r * cols + cc];
        }
}

// thread per (entry slot): interfacesUpper coefficients of the coupled boundary faces <-> intUpper[NB*nc] (boundary-face order)
__global__ void k_ldu_interface(long long nE32, const int* __restrict__ meta, int blk, int F, bool toHost, double* __restrict__ offd,
                                double* __restrict__ intUpper)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nE32) return;
    const int m = meta[i], type = m & 3, fc = m >> 2;
    if (type != ET_COUPLED || fc < F) return;
    const long long ent = i >> 5;
    const int lane = (int)(i & 31);
    int r0, c0;
    const int rows = blockRows(blk, &r0), cols = blockCols(blk, &c0), nc = rows * cols;
    const size_t b = (size_t)(fc - F);
    for (int r = 0; r < rows; r++)
        for (int cc = 0; cc < cols; cc++) {
            double* v = offd + ((size_t)ent * 25 + (r0 + r) * 5 + (c0 + cc)) * 32 + lane;
            if (toHost) intUpper[b * nc + r * cols + cc] = *v;
            else *v = intUpper[b * nc + r * cols + cc];
        }
}

__global__ void k_ldu_diag(int NP, const int* __restrict__ pos2cell, int blk, bool toHost, double* __restrict__ diag, double* __restrict__ out)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    int cell = pos2cell[p];
    if (cell < 0) return;
    int r0, c0;
    const int rows = blockRows(blk, &r0), cols = blockCols(blk, &c0), nc = rows * cols;
    for (int r = 0; r < rows; r++)
        for (int cc = 0; cc < cols; cc++) {
            double* v = diag + (size_t)((r0 + r) * 5 + (c0 + cc)) * NP + p;
            if (toHost) out[(size_t)cell * nc + r * cols + cc] = *v;
            else *v = out[(size_t)cell * nc + r * cols + cc];
        }
}

__global__ void k_rdiag(int NP, const int* __restrict__ pos2cell, const double* __restrict__ diag, double* __restrict__ rD, int* __restrict__ err)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    if (pos2cell[p] < 0) { rD[p] = 0.0; return; }
    double r = ICS_GREAT;
    r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)0 * NP + p]));
    r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)24 * NP + p]));
    r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)6 * NP + p]));
    r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)12 * NP + p]));
    r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)18 * NP + p]));
    rD[p] = r;
    if (r < ICS_VSMALL) atomicOr(err, 1);
}

}  // namespace

int ics_jacobian(icsb200_ctx* c, bool useStoredRdt)
{
    JacArgs a{};
    a.NP = c->NP; a.NB = c->NB; a.F = c->F;
    a.pos2cell = c->d_pos2cell; a.sliceOff = c->d_sliceOff; a.rowNAll = c->d_rowNAll; a.col = c->d_col; a.meta = c->d_meta; a.gfid = c->d_gfid;
    a.bfKind = c->d_bfKind;
    a.geo = c->d_geo; a.dCoupled = c->d_dCoupled; a.C = c->d_C; a.V = c->d_V; a.f = c->d_fields; a.grad = c->d_grad; a.vic = c->d_vic; a.co = c->d_co;
    a.NFG = c->NFG; a.NX = c->NX; a.NPH = c->NPH;
    a.limU = c->sch.limiter_U; a.limT = c->sch.limiter_T; a.localDt = c->sch.local_timestepping;
    a.gamma = c->gamma; a.Cv = c->Cv;
    a.ddtScheme = c->sch.ddt_scheme;
    if (a.ddtScheme != ICSB200_DDT_STEADY) {
        a.rDeltaT = 1.0 / c->sch.delta_t;
        double deltaT0 = (c->timeIndex < 2) ? ICS_GREAT : c->sch.delta_t;
        a.coefft = 1 + c->sch.delta_t / (c->sch.delta_t + deltaT0);
    }
    a.mu = c->mu; a.alphaEff = c->gamma * (c->mu / c->Pr);
    a.offd = c->d_offd; a.diag = c->d_diag; a.rD = c->d_rD; a.rdt = c->d_rdt; a.ddtCoeff = c->d_ddtCoeff;
    a.rdtUniform = 0.0;
    a.useStoredRdt = useStoredRdt ? 1 : 0;
    if (!a.localDt && !useStoredRdt) {
        unsigned long long* d_max = (unsigned long long*)(c->d_scal + 4000);
        CUDA_TRY(c, cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), c->stream));
        {
            LaunchScope ls(c, TM_JAC);
            k_lambda_max<<<gridFor(c->NP, 128), 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_rowNLow, c->d_col, c->d_meta,
                                                                    c->d_gfid, c->d_geo, c->NFG, c->d_fields, c->NX, c->d_mrfFace, d_max);
        }
        double mx = 0.0;
        // gMax over all ranks (setCoAndDeltaT.H:143-146); the maximum is non-negative, so its bit pattern orders like the double
        int rr = ics_allreduce_max_double(c, (double*)d_max, 1);
        if (rr) return rr;
        CUDA_TRY(c, cudaMemcpyAsync(&mx, d_max, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        a.rdtUniform = mx / c->pseudoCoNum;
    }
    {
        LaunchScope ls(c, TM_JAC);
        a.recon = c->d_faceRecon;
        a.mrf = c->d_mrfFace; a.mrfOmega = c->d_mrfOmega;
        a.tr = c->d_transport;
        a.R = c->R; a.TMin = c->sch.T_min;
        const bool reuse = c->reconValid && c->d_faceRecon, fullv = c->mu > 0 && c->sch.viscous_full_jacobian;
        if (fullv) {
            if (reuse) k_jac<true, true><<<gridFor(c->NP, 128), 128, 0, c->stream>>>(a);
            else k_jac<false, true><<<gridFor(c->NP, 128), 128, 0, c->stream>>>(a);
        } else if (reuse) k_jac<true, false><<<gridFor(c->NP, 128), 128, 0, c->stream>>>(a);
        else k_jac<false, false><<<gridFor(c->NP, 128), 128, 0, c->stream>>>(a);
    }
    CUDA_TRY(c, cudaGetLastError());
    c->matrixSet = true;
    c->rDValid = true;
    c->invDValid = false;
    int r = ics_mrf_source(c);
    if (r) return r;
    if (c->hbNO > 1) return ics_hb_diag(c);  // HB.addBlock(J,J) + the shared lusgs diagonal
    return 0;
}

// addMRFSource, source part (convectiveFluxScheme.C:125-128): dVByV(0,0).source() -= (MRFOmega ^ (rho U)) V.  The reference
// applies it to the fresh eqSystem of every outer iteration; here the sources live in d_src from the residual evaluation
// to the solve, so the flag keeps a repeated assemble() from subtracting it twice.
namespace {
__global__ void k_mrf_source(int NP, const int* __restrict__ pos2cell, const double* __restrict__ omega, const double* __restrict__ f, size_t NX,
                             const double* __restrict__ V, double* __restrict__ src, size_t NPH)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    const double ox = omega[p], oy = omega[(size_t)NP + p], oz = omega[2 * (size_t)NP + p];
    const double rho = f[(size_t)Q_RHO * NX + p];
    const double rux = rho * f[(size_t)Q_UX * NX + p], ruy = rho * f[(size_t)Q_UY * NX + p], ruz = rho * f[(size_t)Q_UZ * NX + p];
    const double cx = oy * ruz - oz * ruy, cy = oz * rux - ox * ruz, cz = ox * ruy - oy * rux;
    const double vol = V[p];
    src[NPH + p] -= cx * vol;
    src[2 * NPH + p] -= cy * vol;
    src[3 * NPH + p] -= cz * vol;
}
}  // namespace

int ics_mrf_source(icsb200_ctx* c)
{
    if (!c->d_mrfOmega || c->srcMrfApplied) return 0;
    {
        LaunchScope ls(c, TM_JAC);
        k_mrf_source<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_mrfOmega, c->d_fields, c->NX, c->d_V, c->d_src, c->NPH);
    }
    CUDA_TRY(c, cudaGetLastError());
    c->srcMrfApplied = true;
    return 0;
}

int ics_rdiag(icsb200_ctx* c)
{
    if (c->hbNO > 1) return ics_hb_rdiag(c);
    int* err = (int*)c->d_counter + 40;
    CUDA_TRY(c, cudaMemsetAsync(err, 0, sizeof(int), c->stream));
    {
        LaunchScope ls(c, TM_JAC);
        k_rdiag<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_diag, c->d_rD, err);
    }
    int h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (h) return ics_fail(c, ICSB200_ESINGULAR, "All diagonals of coupledMatrix are zero.");
    c->rDValid = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI
static const int kBlockNc[9] = {1, 1, 1, 1, 3, 3, 3, 3, 9};

extern "C" int icsb200_matrix_get_ldu(icsb200_ctx* c, int block, double* diag, double* upper, double* lower)
{
    if (!c->matrixSet) return ics_fail(c, ICSB200_ESTATE, "matrix_get_ldu: matrix not assembled");
    if (block < 0 || block > 8) return ics_fail(c, ICSB200_EINVAL, "matrix_get_ldu: block id");
    const int nc = kBlockNc[block];
    const size_t nd = (size_t)nc * c->N, nf = (size_t)nc * c->F;
    int r = ics_ensure_stage(c, sizeof(double) * (nd + 2 * nf + 1));
    if (r) return r;
    double *sd = c->d_stage, *su = sd + nd, *sl = su + nf;
    CUDA_TRY(c, cudaMemsetAsync(sd, 0, sizeof(double) * (nd + 2 * nf), c->stream));
    const long long nE32 = c->nEntries * 32;
    {
        LaunchScope ls(c, TM_PERM);
        k_ldu_diag<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, block, true, c->d_diag, sd);
        if (nE32 > 0) k_ldu_offdiag<<<gridFor(nE32, 256), 256, 0, c->stream>>>(nE32, c->d_meta, nullptr, block, c->F, true, c->d_offd, su, sl);
    }
    CUDA_TRY(c, cudaGetLastError());
    if (diag) CUDA_TRY(c, cudaMemcpyAsync(diag, sd, sizeof(double) * nd, cudaMemcpyDeviceToHost, c->stream));
    if (upper && nf) CUDA_TRY(c, cudaMemcpyAsync(upper, su, sizeof(double) * nf, cudaMemcpyDeviceToHost, c->stream));
    if (lower && nf) CUDA_TRY(c, cudaMemcpyAsync(lower, sl, sizeof(double) * nf, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int icsb200_matrix_set_ldu(icsb200_ctx* c, int block, const double* diag, const double* upper, const double* lower)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "matrix_set_ldu: mesh not set");
    if (block < 0 || block > 8 || !diag) return ics_fail(c, ICSB200_EINVAL, "matrix_set_ldu: bad argument");
    if ((upper == nullptr) != (lower == nullptr)) return ics_fail(c, ICSB200_EINVAL, "matrix_set_ldu: upper and lower must come together");
    const int nc = kBlockNc[block];
    const size_t nd = (size_t)nc * c->N, nf = (size_t)nc * c->F;
    int r = ics_ensure_stage(c, sizeof(double) * (nd + 2 * nf + 1));
    if (r) return r;
    double *sd = c->d_stage, *su = sd + nd, *sl = su + nf;
    CUDA_TRY(c, cudaMemcpyAsync(sd, diag, sizeof(double) * nd, cudaMemcpyHostToDevice, c->stream));
    if (upper && nf) {
        CUDA_TRY(c, cudaMemcpyAsync(su, upper, sizeof(double) * nf, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(sl, lower, sizeof(double) * nf, cudaMemcpyHostToDevice, c->stream));
    } else if (nf) {
        CUDA_TRY(c, cudaMemsetAsync(su, 0, sizeof(double) * 2 * nf, c->stream));
    }
    const long long nE32 = c->nEntries * 32;
    {
        LaunchScope ls(c, TM_PERM);
        k_ldu_diag<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->d_pos2cell, block, false, c->d_diag, sd);
        if (nE32 > 0 && nf) k_ldu_offdiag<<<gridFor(nE32, 256), 256, 0, c->stream>>>(nE32, c->d_meta, nullptr, block, c->F, false, c->d_offd, su, sl);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->matrixSet = true;
    c->rDValid = false;
    c->invDValid = false;
    return 0;
}

// interfacesUpper() of one sub-block on the coupled patches (blockFvMatrix.C:248-266; consumed by Amul at blockFvMatrix.C:383-599)
extern "C" int icsb200_matrix_get_interfaces(icsb200_ctx* c, int block, double* intUpper)
{
    if (!c->matrixSet) return ics_fail(c, ICSB200_ESTATE, "matrix_get_interfaces: matrix not assembled");
    if (block < 0 || block > 8 || !intUpper) return ics_fail(c, ICSB200_EINVAL, "matrix_get_interfaces: bad argument");
    const size_t n = (size_t)kBlockNc[block] * c->NB;
    if (n == 0) return 0;
    int r = ics_ensure_stage(c, sizeof(double) * n);
    if (r) return r;
    CUDA_TRY(c, cudaMemsetAsync(c->d_stage, 0, sizeof(double) * n, c->stream));
    const long long nE32 = c->nEntries * 32;
    {
        LaunchScope ls(c, TM_PERM);
        k_ldu_interface<<<gridFor(nE32, 256), 256, 0, c->stream>>>(nE32, c->d_meta, block, c->F, true, c->d_offd, c->d_stage);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(intUpper, c->d_stage, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int icsb200_matrix_set_interfaces(icsb200_ctx* c, int block, const double* intUpper)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "matrix_set_interfaces: mesh not set");
    if (block < 0 || block > 8 || !intUpper) return ics_fail(c, ICSB200_EINVAL, "matrix_set_interfaces: bad argument");
    const size_t n = (size_t)kBlockNc[block] * c->NB;
    if (n == 0) return 0;
    int r = ics_ensure_stage(c, sizeof(double) * n);
    if (r) return r;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, intUpper, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    const long long nE32 = c->nEntries * 32;
    {
        LaunchScope ls(c, TM_PERM);
        k_ldu_interface<<<gridFor(nE32, 256), 256, 0, c->stream>>>(nE32, c->d_meta, block, c->F, false, c->d_offd, c->d_stage);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int icsb200_source_set(icsb200_ctx* c, const double* sRho, const double* sRhoU, const double* sRhoE)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "source_set: mesh not set");
    int r;
    if ((r = ics_upload_cells(c, sRho, 1, c->d_src, c->NPH))) return r;
    if ((r = ics_upload_cells(c, sRhoU, 3, c->d_src + c->NPH, c->NPH))) return r;
    if ((r = ics_upload_cells(c, sRhoE, 1, c->d_src + 4 * (size_t)c->NPH, c->NPH))) return r;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->srcMrfApplied = true;  // caller-provided sources are final (they already carry addMRFSource's term)
    return 0;
}

extern "C" int icsb200_source_get(icsb200_ctx* c, double* sRho, double* sRhoU, double* sRhoE)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "source_get: mesh not set");
    int r = 0;
    if (sRho && (r = ics_download_cells(c, sRho, 1, c->d_src, c->NPH))) return r;
    if (sRhoU && (r = ics_download_cells(c, sRhoU, 3, c->d_src + c->NPH, c->NPH))) return r;
    if (sRhoE && (r = ics_download_cells(c, sRhoE, 1, c->d_src + 4 * (size_t)c->NPH, c->NPH))) return r;
    return 0;
}

// turbulence->muEff() / turbulence->alphaEff() (viscousFluxScheme.C:222-223, residualsUpdate.H:16-43) from the caller's
// turbulence model: cell values [N] and boundary-face values [NB]; coupled boundary entries are ignored (the halo /
// neighbour cell provides the patchNeighbourField).  NULL muEff switches back to the laminar constants.
extern "C" int icsb200_transport_set(icsb200_ctx* c, const double* muEff, const double* muEff_b, const double* alphaEff, const double* alphaEff_b)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "transport_set: mesh not set");
    cudaSetDevice(c->device);
    int r;
    c->fluxValid = false;
    c->matrixSet = false;
    if (!muEff) return devAlloc(c, &c->d_transport, 0);
    if (!muEff_b || !alphaEff || !alphaEff_b) return ics_fail(c, ICSB200_EINVAL, "transport_set: all four arrays or none");
    if (!(c->mu > 0)) return ics_fail(c, ICSB200_ESTATE, "transport_set: the run is inviscid (thermo_set mu = 0)");
    if (!c->d_transport) {
        if ((r = devAlloc(c, &c->d_transport, (size_t)2 * c->NX))) return r;
        CUDA_TRY(c, cudaMemsetAsync(c->d_transport, 0, sizeof(double) * 2 * c->NX, c->stream));
    }
    if ((r = ics_upload_cells(c, muEff, 1, c->d_transport, c->NX))) return r;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));   // the staging buffer is reused by the next upload
    if ((r = ics_upload_cells(c, alphaEff, 1, c->d_transport + c->NX, c->NX))) return r;
    if (c->NB > 0) {
        CUDA_TRY(c, cudaMemcpyAsync(c->d_transport + c->NP + c->NH, muEff_b, sizeof(double) * c->NB, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(c, cudaMemcpyAsync(c->d_transport + c->NX + c->NP + c->NH, alphaEff_b, sizeof(double) * c->NB, cudaMemcpyHostToDevice, c->stream));
    }
    if ((r = ics_halo_fields(c, c->d_transport, c->NX, 2))) return r;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// flux.MRFFaceVelocity() / flux.MRFOmega() as the solver sets them every outer iteration (outerLoop.H:18-21)
extern "C" int icsb200_mrf_set(icsb200_ctx* c, const double* mrf_face_velocity, const double* mrf_omega)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "mrf_set: mesh not set");
    cudaSetDevice(c->device);
    int r;
    if (mrf_face_velocity) {
        std::vector<double> h((size_t)c->NFG);
        for (int g = 0; g < c->NFG; g++) h[g] = mrf_face_velocity[c->h_gf2ref[g]];
        if ((r = devUpload(c, &c->d_mrfFace, h))) return r;
    } else if ((r = devAlloc(c, &c->d_mrfFace, 0))) return r;
    if (mrf_omega) {
        if (!c->d_mrfOmega && (r = devAlloc(c, &c->d_mrfOmega, (size_t)3 * c->NP))) return r;
        CUDA_TRY(c, cudaMemsetAsync(c->d_mrfOmega, 0, sizeof(double) * 3 * c->NP, c->stream));
        if ((r = ics_upload_cells(c, mrf_omega, 3, c->d_mrfOmega, c->NP))) return r;
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    } else if ((r = devAlloc(c, &c->d_mrfOmega, 0))) return r;
    c->fluxValid = false;  // fluxes, sources and matrix of the previous frame velocity are stale
    c->matrixSet = false;
    return 0;
}

extern "C" int icsb200_pseudo_dt(icsb200_ctx* c, double* rPseudoDeltaT, double* pseudoCo)
{
    // setCoAndDeltaT.H: SER update of the pseudo Courant number, then the local pseudo time step.  On the device the
    // time-step part is fused into the Jacobian kernel (same lambda); this entry point runs both and reads the result.
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "pseudo_dt: state not set");
    int r;
    if ((r = ics_pseudo_ser(c))) return r;
    if ((r = ics_gradients(c))) return r;
    if ((r = ics_jacobian(c, false))) return r;
    if (rPseudoDeltaT && (r = ics_download_cells(c, rPseudoDeltaT, 1, c->d_rdt, c->NP))) return r;
    if (pseudoCo && (r = ics_download_cells(c, pseudoCo, 1, c->d_co, c->NP))) return r;
    return 0;
}

extern "C" int icsb200_assemble(icsb200_ctx* c)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "assemble: state not set");
    int r;
    // piecewise use (parity API): the pseudo time step is the one left by the last pseudo_dt() call, as in the
    // reference where createConvectiveJacobian only sees ddtCoeff (outerLoop.H:61-78)
    if ((r = ics_gradients(c))) return r;
    if ((r = ics_copy_prev(c))) return r;
    if ((r = ics_jacobian(c, true))) return r;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}
