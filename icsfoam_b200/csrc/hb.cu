// hb.cu — Harmonic Balance: HBZone source terms, diagonal blocks, shared LU-SGS diagonal and the dense
// block-Jacobi preconditioner of the (2 nO, nO) coupled system of dbnsFullyImplicitHBFoam.
//
// Reference: src/cfdTools/HB/HBZoneTemplates.C:38-92 (addSource), HBZone.C:435-518 (addBlock), HBZone.C:521-651
// (cylindrical momentum source), applications/solvers/dbnsFullyImplicitHBFoam/outerLoop.H:28-30,157-206,
// lusgs.C:50-123 (rDiagCoeff over every diagonal of the global matrix), JacobiSmoother.C:42-203.
//
// B200 design: the nO time instances are not nO meshes and nO^2 block slots, they are ONE mesh of nO disconnected
// copies, so every kernel of the single-instance path (gradients, flux, Jacobian, SpMV, LU-SGS, vector ops) runs
// unchanged on nO x wider levels.  The only inter-instance coupling of the reference's global system is the diagonal
// V*D[J][K] of the (rho,rho), (rhoU,rhoU), (rhoE,rhoE) blocks; here that is a gather through d_hbPeer (the position
// of the same cell in instance K) — no extra matrix storage.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.cuh"

namespace {

struct HBZonePrm {
    int cyl;
    double axisHat[3], centre[3];
};

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// unit radial vector of a cell centre w.r.t. the rotation axis (HBZone.C:563-569)
__device__ __forceinline__ void radialHat(const HBZonePrm& z, const double* __restrict__ C, size_t NPH, int q, double* rHat)
{
    double r[3] = {C[q] - z.centre[0], C[NPH + q] - z.centre[1], C[2 * NPH + q] - z.centre[2]};
    const double ar = dot3(z.axisHat, r);
    for (int d = 0; d < 3; d++) r[d] -= ar * z.axisHat[d];
    const double magr = sqrt(dot3(r, r));
    for (int d = 0; d < 3; d++) rHat[d] = r[d] / magr;
}

// S_J = -V sum_K D[J][K] W_K added to the sources R*V (outerLoop.H:28-30, residualsUpdate.H:72-74)
__global__ void k_hb_source(int NP, int nO, const int* __restrict__ peer, const int* __restrict__ inst, const int* __restrict__ zone,
                            const double* __restrict__ D, const HBZonePrm* __restrict__ zp, const double* __restrict__ V, const double* __restrict__ C,
                            size_t NPH, const double* __restrict__ f, size_t NX, double* __restrict__ src)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    const int J = inst[p];
    if (J < 0) return;
    const int z = zone[p];
    if (z < 0) return;
    const double* Dz = D + ((size_t)z * nO + J) * nO;
    const double vol = V[p];
    const HBZonePrm prm = zp[z];
    double s0 = 0.0, s4 = 0.0, sm[3] = {0.0, 0.0, 0.0};
    for (int K = 0; K < nO; K++) {
        const int q = peer[(size_t)K * NP + p];
        const double vd = vol * Dz[K];
        s0 -= vd * f[(size_t)Q_W0 * NX + q];
        s4 -= vd * f[(size_t)Q_W4 * NX + q];
    }
    if (!prm.cyl) {
        for (int K = 0; K < nO; K++) {
            const int q = peer[(size_t)K * NP + p];
            const double vd = vol * Dz[K];
            for (int d = 0; d < 3; d++) sm[d] -= vd * f[(size_t)(Q_W1 + d) * NX + q];
        }
    } else {
        double sourceCyl[3] = {0.0, 0.0, 0.0};
        for (int K = 0; K < nO; K++) {
            const int q = peer[(size_t)K * NP + p];
            double rHat[3], tHat[3];
            radialHat(prm, C, NPH, q, rHat);
            tHat[0] = prm.axisHat[1] * rHat[2] - prm.axisHat[2] * rHat[1];
            tHat[1] = prm.axisHat[2] * rHat[0] - prm.axisHat[0] * rHat[2];
            tHat[2] = prm.axisHat[0] * rHat[1] - prm.axisHat[1] * rHat[0];
            const double u[3] = {f[(size_t)Q_W1 * NX + q], f[(size_t)Q_W2 * NX + q], f[(size_t)Q_W3 * NX + q]};
            const double UCyl[3] = {dot3(u, rHat), dot3(u, tHat), dot3(u, prm.axisHat)};
            const double vd = vol * Dz[K];
            for (int d = 0; d < 3; d++) sourceCyl[d] += vd * UCyl[d];
        }
        double rHat[3], tHat[3];
        radialHat(prm, C, NPH, p, rHat);
        tHat[0] = prm.axisHat[1] * rHat[2] - prm.axisHat[2] * rHat[1];
        tHat[1] = prm.axisHat[2] * rHat[0] - prm.axisHat[0] * rHat[2];
        tHat[2] = prm.axisHat[0] * rHat[1] - prm.axisHat[1] * rHat[0];
        for (int d = 0; d < 3; d++) {
            const double sourceCart = sourceCyl[0] * rHat[d] + sourceCyl[1] * tHat[d] + sourceCyl[2] * prm.axisHat[d];
            sm[d] -= sourceCart;
        }
    }
    src[p] = src[p] + s0;
    for (int d = 0; d < 3; d++) src[(size_t)(1 + d) * NPH + p] = src[(size_t)(1 + d) * NPH + p] + sm[d];
    src[4 * NPH + p] = src[4 * NPH + p] + s4;
}

// HB.addBlock(eqSystemBlock.dSByS(0,0) / dVByV(0,0) / dSByS(1,1), K, K)  (outerLoop.H:161-163)
__global__ void k_hb_diag(int NP, int nO, const int* __restrict__ inst, const int* __restrict__ zone, const double* __restrict__ D,
                          const double* __restrict__ V, double* __restrict__ diag)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    const int J = inst[p];
    if (J < 0) return;
    const int z = zone[p];
    if (z < 0) return;
    const double vd = V[p] * D[((size_t)z * nO + J) * nO + J];
    diag[(size_t)0 * NP + p] += vd;
    diag[(size_t)6 * NP + p] += vd * 1.0;
    diag[(size_t)12 * NP + p] += vd * 1.0;
    diag[(size_t)18 * NP + p] += vd * 1.0;
    diag[(size_t)24 * NP + p] += vd;
}

// lusgs::lusgs (lusgs.C:50-123) for nScalar = 2 nO, nVector = nO: one coefficient per cell, shared by all instances
__global__ void k_hb_rdiag(int NP, int nO, const int* __restrict__ peer, const int* __restrict__ inst, const double* __restrict__ diag,
                           double* __restrict__ rD, int* __restrict__ err)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    if (inst[p] < 0) { rD[p] = 0.0; return; }
    double r = ICS_GREAT;
    for (int K = 0; K < nO; K++) {
        const int q = peer[(size_t)K * NP + p];
        r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)0 * NP + q]));
        r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)24 * NP + q]));
    }
    for (int K = 0; K < nO; K++) {
        const int q = peer[(size_t)K * NP + p];
        r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)6 * NP + q]));
        r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)12 * NP + q]));
        r = 1.0 / fmax(1.0 / r, fabs(diag[(size_t)18 * NP + q]));
    }
    rD[p] = r;
    if (r < ICS_VSMALL) atomicOr(err, 1);
}

// ---- dense block-Jacobi of the global system: (5 nO)^2 matrix per cell (JacobiSmoother.C:42-101) ----
// variable order: scalars (rho_0, rhoE_0, rho_1, rhoE_1, ...) then vectors (rhoU_0 xyz, rhoU_1 xyz, ...).
// DEVIATION: the reference writes dVByV(v, nv) at column nScalar + nv (+1, +2), which overlaps
// for nVector > 1; the intended nScalar + 3 nv is used.  One thread per instance-0 cell; the matrix lives in global
// scratch, strided by cell so that neighbouring threads coalesce.
__device__ __forceinline__ int blkIndexOfVar(int nO, int v, int* instOut)
{
    // global variable index -> (instance, component of the 5-block (rho, rhoUx, rhoUy, rhoUz, rhoE))
    const int nS = 2 * nO;
    if (v < nS) { *instOut = v >> 1; return (v & 1) ? 4 : 0; }
    const int w = v - nS;
    *instOut = w / 3;
    return 1 + (w % 3);
}

__global__ void k_hb_jacobi_invert(int NC, int NP, int nO, const int* __restrict__ cellPos0, const int* __restrict__ peer, const int* __restrict__ zone,
                                   const double* __restrict__ D, const double* __restrict__ V, const double* __restrict__ diag, double* __restrict__ a,
                                   double* __restrict__ inv, double* __restrict__ work, int* __restrict__ pivw)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= NC) return;
    const int n = 5 * nO;
    const int p0 = cellPos0[c];
    const int z = zone[p0];
#define A_(i, j) a[((size_t)(i) * n + (j)) * NC + c]
#define INV_(i, j) inv[((size_t)(i) * n + (j)) * NC + c]
#define VV_(i) work[(size_t)(i) * NC + c]
#define X_(i) work[(size_t)(n + (i)) * NC + c]
#define PIV_(i) pivw[(size_t)(i) * NC + c]
    for (int i = 0; i < n; i++) {
        int I, ci;
        ci = blkIndexOfVar(nO, i, &I);
        const int pI = peer[(size_t)I * NP + p0];
        for (int j = 0; j < n; j++) {
            int K, cj;
            cj = blkIndexOfVar(nO, j, &K);
            double v;
            if (K == I) v = diag[(size_t)(ci * 5 + cj) * NP + pI];
            else v = (ci == cj && z >= 0) ? V[pI] * D[((size_t)z * nO + I) * nO + K] : 0.0;
            A_(i, j) = v;
        }
    }
    // LUscalarMatrix::inv : Foam::LUDecompose (partial pivoting, implicit scaling) + back substitution column by column
    for (int i = 0; i < n; i++) {
        double largest = 0.0;
        for (int j = 0; j < n; j++) largest = fmax(largest, fabs(A_(i, j)));
        VV_(i) = 1.0 / largest;
    }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) { double sum = A_(i, j); for (int k = 0; k < i; k++) sum -= A_(i, k) * A_(k, j); A_(i, j) = sum; }
        int iMax = 0;
        double largest = 0.0;
        for (int i = j; i < n; i++) {
            double sum = A_(i, j);
            for (int k = 0; k < j; k++) sum -= A_(i, k) * A_(k, j);
            A_(i, j) = sum;
            const double temp = VV_(i) * fabs(sum);
            if (temp >= largest) { largest = temp; iMax = i; }
        }
        PIV_(j) = iMax;
        if (j != iMax) {
            for (int k = 0; k < n; k++) { const double t = A_(iMax, k); A_(iMax, k) = A_(j, k); A_(j, k) = t; }
            VV_(iMax) = VV_(j);
        }
        if (A_(j, j) == 0.0) A_(j, j) = ICS_SMALL;
        if (j != n - 1) { const double rDiag = 1.0 / A_(j, j); for (int i = j + 1; i < n; i++) A_(i, j) *= rDiag; }
    }
    for (int cc = 0; cc < n; cc++) {
        for (int i = 0; i < n; i++) X_(i) = 0.0;
        X_(cc) = 1.0;
        int ii = 0;
        for (int i = 0; i < n; i++) {
            const int ip = PIV_(i);
            double sum = X_(ip);
            X_(ip) = X_(i);
            if (ii != 0) { for (int j = ii - 1; j < i; j++) sum -= A_(i, j) * X_(j); }
            else if (sum != 0.0) ii = i + 1;
            X_(i) = sum;
        }
        for (int i = n - 1; i >= 0; i--) {
            double sum = X_(i);
            for (int j = i + 1; j < n; j++) sum -= A_(i, j) * X_(j);
            X_(i) = sum / A_(i, i);
        }
        for (int i = 0; i < n; i++) INV_(i, cc) = X_(i);
    }
#undef A_
#undef VV_
#undef X_
#undef PIV_
}

// x = D^-1 b for all instances of a cell (JacobiSmoother.C:123-203 with a zero initial guess)
__global__ void k_hb_jacobi_apply(int NC, int NP, int nO, const int* __restrict__ cellPos0, const int* __restrict__ peer, const double* __restrict__ inv,
                                  double* __restrict__ x, size_t NPH, double* __restrict__ work)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= NC) return;
    const int n = 5 * nO;
    const int p0 = cellPos0[c];
    for (int j = 0; j < n; j++) {
        int K;
        const int cj = blkIndexOfVar(nO, j, &K);
        work[(size_t)j * NC + c] = -(0.0 - x[(size_t)cj * NPH + peer[(size_t)K * NP + p0]]);
    }
    for (int i = 0; i < n; i++) {
        double res = 0.0;
        for (int j = 0; j < n; j++) res += INV_(i, j) * work[(size_t)j * NC + c];
        int I;
        const int ci = blkIndexOfVar(nO, i, &I);
        x[(size_t)ci * NPH + peer[(size_t)I * NP + p0]] = res;
    }
#undef INV_
}

}  // namespace

int ics_hb_source(icsb200_ctx* c)
{
    if (c->hbNO <= 1) return 0;
    LaunchScope ls(c, TM_FLUX);
    k_hb_source<<<gridFor(c->NP, 128), 128, 0, c->stream>>>(c->NP, c->hbNO, c->d_hbPeer, c->d_hbInst, c->d_hbZone, c->d_hbD, (const HBZonePrm*)c->d_hbZonePrm,
                                                           c->d_V, c->d_C, c->NPH, c->d_fields, c->NX, c->d_src);
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_hb_diag(icsb200_ctx* c)
{
    if (c->hbNO <= 1) return 0;
    {
        LaunchScope ls(c, TM_JAC);
        k_hb_diag<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->hbNO, c->d_hbInst, c->d_hbZone, c->d_hbD, c->d_V, c->d_diag);
    }
    CUDA_TRY(c, cudaGetLastError());
    c->rDValid = false;
    c->invDValid = false;
    return ics_hb_rdiag(c);
}

int ics_hb_rdiag(icsb200_ctx* c)
{
    int* err = (int*)c->d_counter + 40;
    CUDA_TRY(c, cudaMemsetAsync(err, 0, sizeof(int), c->stream));
    {
        LaunchScope ls(c, TM_JAC);
        k_hb_rdiag<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, c->hbNO, c->d_hbPeer, c->d_hbInst, c->d_diag, c->d_rD, err);
    }
    int h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (h) return ics_fail(c, ICSB200_ESINGULAR, "All diagonals of coupledMatrix are zero.");
    c->rDValid = true;
    return 0;
}

int ics_hb_jacobi(icsb200_ctx* c, double* x)
{
    const int nO = c->hbNO, n = 5 * nO, NC = c->N / nO;
    int r;
    if (!c->d_hbInv) {
        if ((r = devAlloc(c, &c->d_hbInv, (size_t)2 * n * n * NC))) return r;    // inverse + LU workspace
        if ((r = devAlloc(c, &c->d_hbWork, (size_t)3 * n * NC))) return r;        // vv, x (doubles) and pivots (ints, in the last n*NC doubles)
    }
    double* inv = c->d_hbInv;
    double* lu = c->d_hbInv + (size_t)n * n * NC;
    // positions of the instance-0 cells: cell2pos[0 .. NC)
    if (!c->invDValid) {
        LaunchScope ls(c, TM_JACOBI);
        k_hb_jacobi_invert<<<gridFor(NC, 64), 64, 0, c->stream>>>(NC, c->NP, nO, c->d_cell2pos, c->d_hbPeer, c->d_hbZone, c->d_hbD, c->d_V, c->d_diag, lu, inv,
                                                                 c->d_hbWork, (int*)(c->d_hbWork + (size_t)2 * n * NC));
        CUDA_TRY(c, cudaGetLastError());
        c->invDValid = true;
    }
    LaunchScope ls(c, TM_JACOBI);
    k_hb_jacobi_apply<<<gridFor(NC, 64), 64, 0, c->stream>>>(NC, c->NP, nO, c->d_cell2pos, c->d_hbPeer, inv, x, c->NPH, c->d_hbWork);
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------ C ABI
extern "C" int icsb200_hb_set(icsb200_ctx* c, int n_instants, int n_zones, const double* D, const int* zone_of_cell, const int* cyl_coords,
                              const double* rotation_axis, const double* rotation_centre)
{
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "hb_set: mesh not set");
    cudaSetDevice(c->device);
    if (n_instants <= 1) { c->hbNO = 1; return 0; }
    // multi-rank: every rank holds all n_instants instances of its own cells (HB instants are not sharded); halos
    // run per replicated processor patch, residual sums and dot products are all-reduced like in the single-instance solver
    if (n_instants > 16) return ics_fail(c, ICSB200_EINVAL, "hb_set: at most 16 time instances");
    if (n_zones < 1 || !D) return ics_fail(c, ICSB200_EINVAL, "hb_set: no HB zone");
    const int nO = n_instants;
    if (c->N % nO || c->F % nO || c->FT % nO) return ics_fail(c, ICSB200_EINVAL, "hb_set: the mesh is not made of n_instants identical copies");
    const int NC = c->N / nO, FC = c->F / nO;
    for (int K = 1; K < nO; K++)
        for (int f = 0; f < FC; f++)
            if (c->owner[(size_t)K * FC + f] != c->owner[f] + K * NC || c->neighbour[(size_t)K * FC + f] != c->neighbour[f] + K * NC)
                return ics_fail(c, ICSB200_EINVAL, "hb_set: the mesh is not made of n_instants identical, instance-major copies");
    const int NP = c->NP;
    std::vector<int> peer((size_t)nO * NP, -1), inst(NP, -1), zone(NP, -1);
    for (int p = 0; p < NP; p++) {
        const int cell = c->pos2cell[p];
        if (cell < 0) continue;
        const int J = cell / NC, cc = cell % NC;
        inst[p] = J;
        const int z = zone_of_cell ? zone_of_cell[cc] : 0;
        if (z >= n_zones) return ics_fail(c, ICSB200_EINVAL, "hb_set: zone index out of range");
        zone[p] = z < 0 ? -1 : z;
        for (int K = 0; K < nO; K++) peer[(size_t)K * NP + p] = c->cell2pos[(size_t)K * NC + cc];
    }
    std::vector<double> Dv(D, D + (size_t)n_zones * nO * nO);
    std::vector<HBZonePrm> zp(n_zones);
    for (int z = 0; z < n_zones; z++) {
        zp[z].cyl = cyl_coords ? cyl_coords[z] : 0;
        double ax[3] = {0, 0, 1}, ce[3] = {0, 0, 0};
        if (zp[z].cyl) {
            if (!rotation_axis || !rotation_centre) return ics_fail(c, ICSB200_EINVAL, "hb_set: cylCoords needs rotationAxis and rotationCentre");
            for (int d = 0; d < 3; d++) { ax[d] = rotation_axis[3 * z + d]; ce[d] = rotation_centre[3 * z + d]; }
        }
        const double magAxis = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
        for (int d = 0; d < 3; d++) { zp[z].axisHat[d] = ax[d] / magAxis; zp[z].centre[d] = ce[d]; }
    }
    int r = 0;
    r |= devUpload(c, &c->d_hbPeer, peer);
    r |= devUpload(c, &c->d_hbInst, inst);
    r |= devUpload(c, &c->d_hbZone, zone);
    r |= devUpload(c, &c->d_hbD, Dv);
    if (r) return r;
    if (c->d_hbZonePrm) { cudaFree(c->d_hbZonePrm); c->d_hbZonePrm = nullptr; }
    CUDA_TRY(c, cudaMalloc(&c->d_hbZonePrm, sizeof(HBZonePrm) * n_zones));
    CUDA_TRY(c, cudaMemcpy(c->d_hbZonePrm, zp.data(), sizeof(HBZonePrm) * n_zones, cudaMemcpyHostToDevice));
    if (c->d_hbInv) { cudaFree(c->d_hbInv); c->d_hbInv = nullptr; }
    if (c->d_hbWork) { cudaFree(c->d_hbWork); c->d_hbWork = nullptr; }
    c->hbNO = nO;
    c->hbNZones = n_zones;
    c->hbSInit.assign(2 * nO, 0.0); c->hbVInit.assign(3 * nO, 0.0); c->hbSFinal.assign(2 * nO, 0.0); c->hbVFinal.assign(3 * nO, 0.0);
    c->hbSInitPrev.clear(); c->hbVInitPrev.clear();
    c->matrixSet = false;
    c->fluxValid = false;
    c->rDValid = c->invDValid = false;
    return 0;
}

extern "C" int icsb200_hb_residuals_get(icsb200_ctx* c, double* s_init, double* v_init, double* s_final, double* v_final)
{
    if (c->hbNO <= 1) return ics_fail(c, ICSB200_ESTATE, "hb_residuals_get: Harmonic Balance is not set");
    const int nO = c->hbNO;
    if (s_init) std::memcpy(s_init, c->hbSInit.data(), sizeof(double) * 2 * nO);
    if (v_init) std::memcpy(v_init, c->hbVInit.data(), sizeof(double) * 3 * nO);
    if (s_final) std::memcpy(s_final, c->hbSFinal.data(), sizeof(double) * 2 * nO);
    if (v_final) std::memcpy(v_final, c->hbVFinal.data(), sizeof(double) * 3 * nO);
    return 0;
}
