// solver.cu — coupledMatrix Krylov solve on the device: block SpMV, LU-SGS / block-Jacobi, restarted GMRES.
//
// Replaces coupledMatrix::matrixMul (coupledMatrix.C:66-123) = 9 x blockFvMatrix::Amul (blockFvMatrix.C:329-601),
// lusgs (lusgs.C:50-382), Jacobi / JacobiSmoother (Jacobi.C:55-132, JacobiSmoother.C:42-203),
// gmres::solveDelta 6-arg (gmres.C:772-1110) with solver::stop (coupledMatrixSolver.C:198-221) and the halo /
// reduction traffic of Pstream (blockFvMatrixUpdateMatrixInterfaces.C:33-185, gmres.C:939-1008).
//
//  * SpMV: one thread per block row of the sliced-ELL storage; 25 coalesced streaming loads per 5x5 block, x gathered
//    through L1/L2.  The per-row accumulation keeps the reference's three partial sums per equation (rho-, rhoE-,
//    rhoU-columns) and the ascending-face order, so the product is bit-identical to the 9-sub-block Amul chain.
//  * LU-SGS: exact Gauss-Seidel order of the reference by level scheduling (levels = longest path in the
//    owner<neighbour DAG; cells are stored level by level), one persistent cooperative kernel for both sweeps with a
//    grid barrier between levels; rows pull their lower (forward) / upper (reverse) neighbours in the order the
//    reference pushes them.
//  * GMRES: modified Gram-Schmidt with every scalar (H, Givens, beta) resident on the device; each MGS step is one
//    fused pass (w -= h_j v_j, then the next dot product / norm from registers); reductions are two-stage
//    warp-shuffle trees with a fixed combination order (run-to-run reproducible); one host read-back per restart.
#include <cooperative_groups.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ halo
__global__ void k_pack(int n, int nArrays, const int* __restrict__ sendPos, const double* __restrict__ base, size_t stride, double* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int p = sendPos[i];
    for (int k = 0; k < nArrays; k++) out[(size_t)k * n + i] = base[k * stride + p];
}
__global__ void k_unpack(int n, int nArrays, int slot0, const double* __restrict__ in, double* __restrict__ base, size_t stride)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < nArrays; k++) base[k * stride + slot0 + i] = in[(size_t)k * n + i];
}

}  // namespace

// exchange the processor-patch halo of nArrays SoA arrays (patchNeighbourField of every coupled field at once)
// cyclicAMIFvPatchField::patchNeighbourField (cyclicAMIFvPatchField.C:146-209): halo slot = sum_k w_k phi[srcPos_k], summed in
// the order of the AMI address list (AMIInterpolation::interpolateToSource with plusEqOp: result = 0; result += w*phi)
__global__ void k_ami_gather(int n, int nArrays, int slot0, const int* __restrict__ start, const int* __restrict__ srcPos, const double* __restrict__ w,
                             double* __restrict__ base, size_t stride)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * nArrays) return;
    const int a = t / n, i = t - a * n;
    double acc = 0.0;
    for (int k = start[i]; k < start[i + 1]; k++) acc += w[k] * base[(size_t)a * stride + srcPos[k]];
    base[(size_t)a * stride + slot0 + i] = acc;
}

// rotational cyclicAMI: transform(forwardT, interpolated value) of the vector triples, in place on the halo slots
__global__ void k_rot_inplace(int n, int nArrays, int slot0, double t0, double t1, double t2, double t3, double t4, double t5, double t6, double t7,
                              double t8, unsigned vecMask, double* __restrict__ base, size_t stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int dst = slot0 + i;
    for (int a = 0; a + 2 < nArrays; a++) {
        if (!(vecMask >> a & 1u)) continue;
        const double v0 = base[(size_t)a * stride + dst], v1 = base[(size_t)(a + 1) * stride + dst], v2 = base[(size_t)(a + 2) * stride + dst];
        base[(size_t)a * stride + dst] = t0 * v0 + t1 * v1 + t2 * v2;
        base[(size_t)(a + 1) * stride + dst] = t3 * v0 + t4 * v1 + t5 * v2;
        base[(size_t)(a + 2) * stride + dst] = t6 * v0 + t7 * v1 + t8 * v2;
        a += 2;
    }
}

static int amiGather(icsb200_ctx* c, double* base, size_t stride, int nArrays, unsigned vecMask)
{
    if (c->amis.empty()) return 0;
    LaunchScope ls(c, TM_HALO);
    for (auto& am : c->amis) {
        k_ami_gather<<<gridFor((long long)am.size * nArrays, 128), 128, 0, c->stream>>>(am.size, nArrays, c->NP + am.haloStart, am.d_start, am.d_srcPos, am.d_w, base, stride);
        if (am.rot && vecMask) {
            k_rot_inplace<<<gridFor(am.size, 128), 128, 0, c->stream>>>(am.size, nArrays, c->NP + am.haloStart, am.T[0], am.T[1], am.T[2], am.T[3], am.T[4],
                                                                      am.T[5], am.T[6], am.T[7], am.T[8], vecMask, base, stride);
            c->launches++;
        }
    }
    c->launches += (long long)c->amis.size() - 1;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

// cyclicFvPatchField::patchNeighbourField with doTransform() (originalOFFiles/constraintFvPatchFields/cyclic/
// cyclicFvPatchField.C:130-190): halo slot = transform(forwardT, neighbour cell value) — the arrays flagged in vecMask are
// the x components of vector triples (U, gradients, the rhoU part of a solver vector) and are rotated, scalars are copied.
// The reference's scalar fields "U.component(i)" take component i of the rotated cell velocity: the same numbers.
// Phase-lag pairs (phaseLagCyclicFvPatchField.C:160-398): the arrays flagged in lagMask take sum_J w_J * (value of time instance J at
// the neighbour cell), J ascending from zero, before the rotation; the others the plain neighbour value.
__global__ void k_rot_gather(int n, int nArrays, int slot0, const int* __restrict__ srcPos, double t0, double t1, double t2, double t3, double t4,
                             double t5, double t6, double t7, double t8, unsigned vecMask, unsigned lagMask, int nLag, const int* __restrict__ lagSrc,
                             LagWeights lw, double* __restrict__ base, size_t stride)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int src = srcPos[i], dst = slot0 + i;
    auto fetch = [&](int a) {
        if (nLag == 0 || !(lagMask >> a & 1u)) return base[(size_t)a * stride + src];
        double acc = 0.0;
        for (int J = 0; J < nLag; J++) acc += lw.w[J] * base[(size_t)a * stride + lagSrc[(size_t)J * n + i]];
        return acc;
    };
    for (int a = 0; a < nArrays;) {
        if (a + 2 < nArrays && (vecMask >> a & 1u)) {
            const double v0 = fetch(a), v1 = fetch(a + 1), v2 = fetch(a + 2);
            base[(size_t)a * stride + dst] = t0 * v0 + t1 * v1 + t2 * v2;
            base[(size_t)(a + 1) * stride + dst] = t3 * v0 + t4 * v1 + t5 * v2;
            base[(size_t)(a + 2) * stride + dst] = t6 * v0 + t7 * v1 + t8 * v2;
            a += 3;
        } else {
            base[(size_t)a * stride + dst] = fetch(a);
            a += 1;
        }
    }
}

static int rotGather(icsb200_ctx* c, double* base, size_t stride, int nArrays, unsigned vecMask, unsigned lagMask)
{
    if (c->rots.empty()) return 0;
    LaunchScope ls(c, TM_HALO);
    for (auto& ro : c->rots)
        k_rot_gather<<<gridFor(ro.size, 128), 128, 0, c->stream>>>(ro.size, nArrays, c->NP + ro.haloStart, ro.d_srcPos, ro.T[0], ro.T[1], ro.T[2], ro.T[3],
                                                                  ro.T[4], ro.T[5], ro.T[6], ro.T[7], ro.T[8], ro.rotate ? vecMask : 0u, lagMask, ro.nLag,
                                                                  ro.d_lagSrc, ro.lagW, base, stride);
    c->launches += (long long)c->rots.size() - 1;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_halo_fields(icsb200_ctx* c, double* base, size_t stride, int nArrays, unsigned vecMask, unsigned lagMask)
{
    if (c->NH == 0) return 0;
    {
        int r = amiGather(c, base, stride, nArrays, vecMask);
        if (r) return r;
        if ((r = rotGather(c, base, stride, nArrays, vecMask, lagMask))) return r;
    }
    if (c->procs.empty()) return 0;
    if (nArrays > 40) return ics_fail(c, ICSB200_EINVAL, "halo: too many arrays");
    LaunchScope ls(c, TM_HALO);
    for (auto& pp : c->procs)
        k_pack<<<gridFor(pp.size, 128), 128, 0, c->stream>>>(pp.size, nArrays, pp.d_sendPos, base, stride, c->d_sendBuf + (size_t)pp.haloStart * nArrays);
    ncclGroupStart();
    for (auto& pp : c->procs) {
        ncclSend(c->d_sendBuf + (size_t)pp.haloStart * nArrays, (size_t)pp.size * nArrays, ncclDouble, pp.nbrRank, (ncclComm_t)c->nccl, c->stream);
        ncclRecv(c->d_recvBuf + (size_t)pp.haloStart * nArrays, (size_t)pp.size * nArrays, ncclDouble, pp.nbrRank, (ncclComm_t)c->nccl, c->stream);
    }
    ncclGroupEnd();
    for (auto& pp : c->procs)
        k_unpack<<<gridFor(pp.size, 128), 128, 0, c->stream>>>(pp.size, nArrays, c->NP + pp.haloStart, c->d_recvBuf + (size_t)pp.haloStart * nArrays, base, stride);
    c->launches += 2 * (long long)c->procs.size() - 1;
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

static int allreduceSum(icsb200_ctx* c, double* d, int n)
{
    if (c->nRanks == 1) return 0;
    if (ncclAllReduce(d, d, n, ncclDouble, ncclSum, (ncclComm_t)c->nccl, c->stream) != ncclSuccess) return ics_fail(c, ICSB200_ECUDA, "ncclAllReduce failed");
    return 0;
}

int ics_allreduce_max_double(icsb200_ctx* c, double* d, int n)
{
    if (c->nRanks == 1) return 0;
    if (ncclAllReduce(d, d, n, ncclDouble, ncclMax, (ncclComm_t)c->nccl, c->stream) != ncclSuccess) return ics_fail(c, ICSB200_ECUDA, "ncclAllReduce failed");
    return 0;
}

int ics_allreduce_max_int(icsb200_ctx* c, int* d, int n)
{
    if (c->nRanks == 1) return 0;
    if (ncclAllReduce(d, d, n, ncclInt, ncclMax, (ncclComm_t)c->nccl, c->stream) != ncclSuccess) return ics_fail(c, ICSB200_ECUDA, "ncclAllReduce failed");
    return 0;
}

namespace {

// ------------------------------------------------------------------------------------------------ SpMV
// HB (template flag): the rows of instance J also receive the diagonal-only blocks dSByS(2J,2K), dSByS(2J+1,2K+1), dVByV(J,K)
// = V D[J][K] of the other instances, accumulated at the place coupledMatrix::matrixMul's (i, j) loops visit them
// (coupledMatrix.C:66-123; dbnsFullyImplicitHBFoam/outerLoop.H:157-204).
template <bool RESID, bool HB>
__global__ void __launch_bounds__(128)
k_spmv(int NP, const int* __restrict__ pos2cell, const int* __restrict__ sliceOff, const int* __restrict__ rowNAll, const int* __restrict__ col,
       const double* __restrict__ offd, const double* __restrict__ diag, const double* __restrict__ x, size_t NPH, const double* __restrict__ b,
       double* __restrict__ y, HBSpmv hb)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    const int lane = p & 31;
    const size_t base = (size_t)sliceOff[p >> 5];
    double xo[5];
#pragma unroll
    for (int k = 0; k < 5; k++) xo[k] = x[k * NPH + p];
    double aR[5], aE[5], aU[5];
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const double d0 = diag[(size_t)(r * 5 + 0) * NP + p], d1 = diag[(size_t)(r * 5 + 1) * NP + p], d2 = diag[(size_t)(r * 5 + 2) * NP + p],
                     d3 = diag[(size_t)(r * 5 + 3) * NP + p], d4 = diag[(size_t)(r * 5 + 4) * NP + p];
        aR[r] = d0 * xo[0];
        aE[r] = d4 * xo[4];
        aU[r] = d1 * xo[1] + d2 * xo[2] + d3 * xo[3];
    }
    const int nAll = rowNAll[p];
    for (int j = 0; j < nAll; j++) {
        const size_t e = (base + j) * 32 + lane;
        const int c = col[e];
        if (c >= (int)NPH) continue;  // physical boundary entry: no coefficient
        const double* blk = offd + ((base + j) * 25) * 32 + lane;
        double xc[5];
#pragma unroll
        for (int k = 0; k < 5; k++) xc[k] = x[k * NPH + c];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const double b0 = __ldcs(blk + (size_t)(r * 5 + 0) * 32), b1 = __ldcs(blk + (size_t)(r * 5 + 1) * 32), b2 = __ldcs(blk + (size_t)(r * 5 + 2) * 32),
                         b3 = __ldcs(blk + (size_t)(r * 5 + 3) * 32), b4 = __ldcs(blk + (size_t)(r * 5 + 4) * 32);
            aR[r] += b0 * xc[0];
            aE[r] += b4 * xc[4];
            aU[r] += b1 * xc[1] + b2 * xc[2] + b3 * xc[3];
        }
    }
    if (HB) {
        const int J = hb.inst[p], z = hb.zone[p];
        const double vol = hb.V[p];
        const double* Dz = hb.D + ((size_t)(z < 0 ? 0 : z) * hb.nO + J) * hb.nO;
        double v[5];
        // scalar rows: SS blocks of instance 0..nO-1 in order, then the own SV block
        double a0 = 0.0, a4 = 0.0;
        for (int K = 0; K < hb.nO; K++) {
            if (K == J) { a0 += aR[0]; a0 += aE[0]; a4 += aR[4]; a4 += aE[4]; }
            else if (z >= 0) {
                const int q = hb.peer[(size_t)K * NP + p];
                const double sK = vol * Dz[K];
                a0 += sK * x[q];
                a4 += sK * x[4 * NPH + q];
            }
        }
        v[0] = a0 + aU[0];
        v[4] = a4 + aU[4];
        // vector rows: own VS blocks, then VV blocks of instance 0..nO-1 in order
        for (int r = 1; r < 4; r++) {
            double au = 0.0;
            au += aR[r];
            au += aE[r];
            for (int K = 0; K < hb.nO; K++) {
                if (K == J) au += aU[r];
                else if (z >= 0) au += (vol * Dz[K]) * x[(size_t)r * NPH + hb.peer[(size_t)K * NP + p]];
            }
            v[r] = au;
        }
#pragma unroll
        for (int r = 0; r < 5; r++) y[r * NPH + p] = RESID ? (b[r * NPH + p] - v[r]) : v[r];
        return;
    }
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const double v = (aR[r] + aE[r]) + aU[r];
        y[r * NPH + p] = RESID ? (b[r * NPH + p] - v) : v;
    }
}

// ------------------------------------------------------------------------------------------------ LU-SGS
// Exact Gauss-Seidel order of lusgs.C:220-382 without barriers and without flags.
// Positions are sorted by forward level, so every lower neighbour of a row lives in a LOWER slice (32 rows = one warp)
// and every upper neighbour in a HIGHER slice.  Warps take slices round-robin in sweep order (ascending for the
// forward sweep, descending for the reverse sweep).  The sweep outputs go to two buffers pre-filled with a sentinel
// bit pattern (a quiet NaN payload no arithmetic can produce): a producer simply stores its 5 values, a consumer
// polls the neighbour's values until none is the sentinel — the data is its own flag (8-byte stores are atomic), so a
// dependency hop costs one store + one L2 round trip, with no fence.  All warps are co-resident (cooperative launch)
// and each walks its slices in dependency order, hence no deadlock; independent levels overlap freely, so a sweep
// costs max(bandwidth, longest dependency chain x hop latency) instead of levels x grid-barrier.
struct LusgsArgs {
    int nSlices;
    const int *sliceOff, *rowNLow, *rowNInt, *col;
    const double *offd, *rD;
    double* x;         // in: right-hand side, out: preconditioned vector
    double *y, *z;     // forward / reverse sweep values, sentinel-filled before launch
    size_t NPH;
    int* err;
    unsigned int busySpins, sleepNs;
    long long* trace;  // optional [nSlices][4] forward-sweep time stamps (ns): start, prefetch issued, deps ready, stored
    int *hintF, *hintR;  // per-slice "probably published" epochs: cheap to poll; correctness comes from the sentinel check
    int epoch;
};

constexpr unsigned long long LUSGS_SENTINEL = 0xFFF8DEADBEEF0B1DULL;

__device__ __forceinline__ double ldPoll(const double* p)
{
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool isSentinel(double v) { return (unsigned long long)__double_as_longlong(v) == LUSGS_SENTINEL; }

__device__ __forceinline__ int ldHint(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// wait until the 5 components of position q in buffer buf have been published: spin on the slice hint (1 sector per
// warp poll), then read the values and make sure none is still the sentinel
__device__ __forceinline__ void pollVec(const double* buf, const int* hint, int epoch, size_t NPH, int q, double* out, int* err, unsigned int busySpins,
                                        unsigned int sleepNs)
{
    unsigned int spins = 0;
    const int* h = hint + (q >> 5);
    while (true) {
        if (ldHint(h) == epoch) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 5; k++) { out[k] = ldPoll(buf + k * NPH + q); ok &= !isSentinel(out[k]); }
            if (ok) return;
        }
        if (++spins > (1u << 24)) { *err = 1; return; }  // never hang the device on a broken schedule
        if (spins > busySpins) __nanosleep(sleepNs);
    }
}

// one neighbour contribution with the block already in registers
__device__ __forceinline__ void lusgsSubReg(double* xr, const double* B, const double* dl)
{
#pragma unroll
    for (int r = 0; r < 5; r++) {
        // sub-block order of lusgs.C:240-303: S.S (rho col, rhoE col) then S.V / V.S then V.V, one subtraction each
        xr[r] -= B[r * 5 + 0] * dl[0];
        xr[r] -= B[r * 5 + 4] * dl[4];
        xr[r] -= B[r * 5 + 1] * dl[1] + B[r * 5 + 2] * dl[2] + B[r * 5 + 3] * dl[3];
    }
}

constexpr int LCH = 3;  // neighbours handled per chunk (hex cells have at most 3 lower and 3 upper neighbours)

// entries [jBeg, jEnd) of row p (ascending if FWD, descending otherwise): everything that does not depend on the sweep
// values (column, rD, 5x5 block) is fetched BEFORE polling, so the dependent part of a hop is poll -> 75 flops -> store
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

template <bool FWD>
__device__ __forceinline__ void lusgsRow(const LusgsArgs& a, int lane, size_t base, int jBeg, int jEnd, double* xr, long long* tr)
{
    const int n = jEnd - jBeg;
    const double* buf = FWD ? a.y : a.z;
    const int* hint = FWD ? a.hintF : a.hintR;
    for (int c0 = 0; c0 < n; c0 += LCH) {
        double B[LCH][25], scale[LCH], dl[LCH][5];
        int q[LCH];
        unsigned int pend = 0;
#pragma unroll
        for (int t = 0; t < LCH; t++) {
            q[t] = 0;
            if (c0 + t < n) {
                const int j = FWD ? (jBeg + c0 + t) : (jEnd - 1 - c0 - t);
                q[t] = a.col[(base + j) * 32 + lane];
                scale[t] = FWD ? a.rD[q[t]] : 1.0;
                const double* blk = a.offd + ((base + j) * 25) * 32 + lane;
#pragma unroll
                for (int k = 0; k < 25; k++) B[t][k] = __ldcs(blk + (size_t)k * 32);
                pend |= 1u << t;
            }
        }
        const unsigned int want = pend;
        if (tr && lane == 0 && c0 == 0) tr[1] = gtime();
        // wait for all neighbours of the chunk at once: hint loads in flight together, then the value loads together
        unsigned int spins = 0;
        while (pend) {
            int hv[LCH];
#pragma unroll
            for (int t = 0; t < LCH; t++) hv[t] = (pend >> t & 1u) ? ldHint(hint + (q[t] >> 5)) : 0;
            unsigned int ready = 0;
#pragma unroll
            for (int t = 0; t < LCH; t++)
                if ((pend >> t & 1u) && hv[t] == a.epoch) {
                    ready |= 1u << t;
#pragma unroll
                    for (int k = 0; k < 5; k++) dl[t][k] = ldPoll(buf + k * a.NPH + q[t]);
                }
#pragma unroll
            for (int t = 0; t < LCH; t++)
                if (ready >> t & 1u) {
                    bool ok = true;
#pragma unroll
                    for (int k = 0; k < 5; k++) ok &= !isSentinel(dl[t][k]);
                    if (ok) pend &= ~(1u << t);
                }
            if (pend) {
                if (++spins > (1u << 24)) { *a.err = 1; break; }  // never hang the device on a broken schedule
                if (spins > a.busySpins) __nanosleep(a.sleepNs);
            }
        }
        if (tr && c0 == 0) atomicMax((unsigned long long*)(tr + 2), (unsigned long long)gtime());
#pragma unroll
        for (int t = 0; t < LCH; t++)
            if (want >> t & 1u) {
                if (FWD) {
#pragma unroll
                    for (int k = 0; k < 5; k++) dl[t][k] = scale[t] * dl[t][k];  // dW*_q = rD_q x_q (lusgs.C:194-216)
                }
                lusgsSubReg(xr, B[t], dl[t]);
            }
    }
}

__global__ void __launch_bounds__(256, 1)
k_lusgs(LusgsArgs a)
{
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, W = (gridDim.x * blockDim.x) >> 5;
    // forward sweep: D dW* = R - L dW*  (y keeps the un-scaled running value, lusgs.C:233-237)
    for (int s = gw; s < a.nSlices; s += W) {
        const int p = s * 32 + lane;
        const size_t base = (size_t)a.sliceOff[s];
        const int nLow = a.rowNLow[p];
        double xr[5];
#pragma unroll
        for (int k = 0; k < 5; k++) xr[k] = a.x[k * a.NPH + p];
        long long* tr = a.trace ? a.trace + 4 * (size_t)s : nullptr;
        if (tr && lane == 0) tr[0] = gtime();
        lusgsRow<true>(a, lane, base, 0, nLow, xr, tr);
#pragma unroll
        for (int k = 0; k < 5; k++) __stcg(a.y + k * a.NPH + p, xr[k]);
        __syncwarp();
        if (lane == 0) __stcg(a.hintF + s, a.epoch);
        if (tr && lane == 0) tr[3] = gtime();
    }
    // reverse sweep: dW = rD (D dW* - U dW), upper neighbours in descending order (lusgs.C:307-381)
    for (int t = gw; t < a.nSlices; t += W) {
        const int s = a.nSlices - 1 - t;
        const int p = s * 32 + lane;
        const size_t base = (size_t)a.sliceOff[s];
        const int nLow = a.rowNLow[p], nInt = a.rowNInt[p];
        const double rd = a.rD[p];
        double xr[5];
        pollVec(a.y, a.hintF, a.epoch, a.NPH, p, xr, a.err, a.busySpins, a.sleepNs);  // own forward value (possibly produced by another warp)
        lusgsRow<false>(a, lane, base, nLow, nInt, xr, nullptr);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double v = rd * xr[k];
            __stcg(a.z + k * a.NPH + p, v);
            a.x[k * a.NPH + p] = v;
        }
        __syncwarp();
        if (lane == 0) __stcg(a.hintR + s, a.epoch);
    }
}

__device__ __forceinline__ void subBlockGlobal(double* xr, const double* __restrict__ blk, const double* dl);

// ---------------- level pipeline with a TMA producer warp ----------------
// Same schedule and protocol as k_lusgs, but the 5x5 blocks of a slice are no longer prefetched into the consumer's
// registers: one producer warp per CTA streams them with cp.async.bulk (TMA, 1-D bulk copies completing on mbarriers)
// into a shared-memory ring many slices ahead of the consumers.  The consumer warps' dependent path is then only
// hint/value polls (L2) + shared-memory reads + 75 flops + stores: no HBM latency can land on the dependency chain,
// and the bulk traffic does not share the LSU queue with the polls.
constexpr int TMA_NC = 8;        // consumer warps per CTA
constexpr int TMA_SE = 3;        // entries (5x5 block columns) staged per slice and sweep
constexpr int TMA_STAGES = 10;   // ring depth
constexpr int TMA_STAGE_DOUBLES = TMA_SE * 25 * 32;

struct TmaArgs {
    int nSlices;
    const int *sliceOff, *rowNLow, *rowNInt, *col;
    const int *sliceFwdHi, *sliceRevLo, *sliceRevHi;  // per slice: staged entry range of each sweep
    const double *offd, *rD;
    double *x, *y, *z;
    size_t NPH;
    int *hintF, *hintR;
    int epoch;
    int* err;
};

__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ bool mbarTryWait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbarTestWait(unsigned long long* bar, unsigned parity)  // non-blocking
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbarWait(unsigned long long* bar, unsigned parity, int* err)
{
    unsigned int spins = 0;
    while (!mbarTryWait(bar, parity)) {
        if (++spins > (1u << 26)) { *err = 2; return false; }
    }
    return true;
}
__device__ __forceinline__ void tmaLoad1D(void* dstSmem, const void* srcGlobal, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal),
                 "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

// one neighbour contribution with the block in shared memory: element (k, lane) at blk[k*32 + lane]
__device__ __forceinline__ void lusgsSubSmem(double* xr, const double* blk, const double* dl)
{
#pragma unroll
    for (int r = 0; r < 5; r++) {
        xr[r] -= blk[(r * 5 + 0) * 32] * dl[0];
        xr[r] -= blk[(r * 5 + 4) * 32] * dl[4];
        xr[r] -= blk[(r * 5 + 1) * 32] * dl[1] + blk[(r * 5 + 2) * 32] * dl[2] + blk[(r * 5 + 3) * 32] * dl[3];
    }
}

__global__ void __launch_bounds__((TMA_NC + 1) * 32, 1)
k_lusgs_tma(TmaArgs a)
{
    extern __shared__ __align__(128) unsigned char smemRaw[];
    double* ring = (double*)smemRaw;                                                   // [TMA_STAGES][TMA_STAGE_DOUBLES]
    unsigned long long* full = (unsigned long long*)(ring + (size_t)TMA_STAGES * TMA_STAGE_DOUBLES);  // [TMA_STAGES]
    unsigned long long* empty = full + TMA_STAGES;
    volatile int* consumed = (volatile int*)(empty + TMA_STAGES);  // [TMA_STAGES] rounds consumed per stage (guards against parity aliasing)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x, b = blockIdx.x;
    // my items: forward slices b, b+G, ... then reverse slices nSlices-1-b, nSlices-1-b-G, ...
    const int nMine = (a.nSlices > b) ? (a.nSlices - b + G - 1) / G : 0;
    const int nItems = 2 * nMine;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; s++) { mbarInit(full + s, 1); mbarInit(empty + s, 1); consumed[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto itemSlice = [&](int i, bool& fwd) { fwd = i < nMine; return fwd ? (b + i * G) : (a.nSlices - 1 - (b + (i - nMine) * G)); };

    if (warp == TMA_NC) {
        // ---------------- producer warp (one elected lane issues the bulk copies) ----------------
        if (lane == 0) {
            for (int i = 0; i < nItems; i++) {
                const int st = i % TMA_STAGES;
                if (i >= TMA_STAGES && !mbarWait(empty + st, ((i / TMA_STAGES) + 1) & 1, a.err)) break;
                bool fwd;
                const int s = itemSlice(i, fwd);
                const int lo = fwd ? 0 : a.sliceRevLo[s];
                const int hi = min(fwd ? a.sliceFwdHi[s] : a.sliceRevHi[s], lo + TMA_SE);
                const unsigned bytes = (hi > lo) ? (unsigned)(hi - lo) * 25 * 32 * 8 : 0;
                if (bytes) {
                    mbarExpectTx(full + st, bytes);
                    tmaLoad1D(ring + (size_t)st * TMA_STAGE_DOUBLES, a.offd + ((size_t)a.sliceOff[s] + lo) * 25 * 32, bytes, full + st);
                } else {
                    mbarArrive(full + st);
                }
            }
        }
        return;
    }
    // ---------------- consumer warps ----------------
    for (int i = warp; i < nItems; i += TMA_NC) {
        const int st = i % TMA_STAGES;
        bool fwd;
        const int s = itemSlice(i, fwd);
        const int p = s * 32 + lane;
        const size_t base = (size_t)a.sliceOff[s];
        const int nLow = a.rowNLow[p], nInt = a.rowNInt[p];
        const int stageLo = fwd ? 0 : a.sliceRevLo[s];
        const int jBeg = fwd ? 0 : nLow, jEnd = fwd ? nLow : nInt;
        const int n = jEnd - jBeg;
        const double* buf = fwd ? a.y : a.z;
        const int* hint = fwd ? a.hintF : a.hintR;
        double xr[5];
        // columns / scales of the first LCH neighbours (tiny, L2)
        int q[LCH];
        double sc[LCH];
#pragma unroll
        for (int t = 0; t < LCH; t++) {
            q[t] = -1;
            if (t < n) {
                const int j = fwd ? (jBeg + t) : (jEnd - 1 - t);
                q[t] = a.col[(base + j) * 32 + lane];
                sc[t] = fwd ? a.rD[q[t]] : 1.0;
            }
        }
        const double rd = fwd ? 1.0 : a.rD[p];
        if (fwd) {
#pragma unroll
            for (int k = 0; k < 5; k++) xr[k] = a.x[k * a.NPH + p];
        } else {
            pollVec(a.y, a.hintF, a.epoch, a.NPH, p, xr, a.err, 64u, 100u);  // own forward value (possibly from another warp)
        }
        // the staged blocks of this item (issued long ago by the producer).  First make sure the previous round of this
        // stage has been consumed: an mbarrier parity wait is only meaningful for the current or the preceding phase
        {
            const int round = i / TMA_STAGES;
            unsigned int spins = 0;
            while (consumed[st] != round) { if (++spins > (1u << 28)) { *a.err = 3; break; } }
        }
        mbarWait(full + st, (i / TMA_STAGES) & 1, a.err);
        const double* stage = ring + (size_t)st * TMA_STAGE_DOUBLES;
        for (int c0 = 0; c0 < n; c0 += LCH) {
            if (c0 > 0) {
#pragma unroll
                for (int t = 0; t < LCH; t++) {
                    q[t] = -1;
                    if (c0 + t < n) {
                        const int j = fwd ? (jBeg + c0 + t) : (jEnd - 1 - c0 - t);
                        q[t] = a.col[(base + j) * 32 + lane];
                        sc[t] = fwd ? a.rD[q[t]] : 1.0;
                    }
                }
            }
            unsigned int pend = 0;
#pragma unroll
            for (int t = 0; t < LCH; t++) if (q[t] >= 0) pend |= 1u << t;
            const unsigned int want = pend;
            double dl[LCH][5];
            unsigned int spins = 0;
            while (pend) {
                int hv[LCH];
#pragma unroll
                for (int t = 0; t < LCH; t++) hv[t] = (pend >> t & 1u) ? ldHint(hint + (q[t] >> 5)) : 0;
                unsigned int ready = 0;
#pragma unroll
                for (int t = 0; t < LCH; t++)
                    if ((pend >> t & 1u) && hv[t] == a.epoch) {
                        ready |= 1u << t;
#pragma unroll
                        for (int k = 0; k < 5; k++) dl[t][k] = ldPoll(buf + k * a.NPH + q[t]);
                    }
#pragma unroll
                for (int t = 0; t < LCH; t++)
                    if (ready >> t & 1u) {
                        bool ok = true;
#pragma unroll
                        for (int k = 0; k < 5; k++) ok &= !isSentinel(dl[t][k]);
                        if (ok) pend &= ~(1u << t);
                    }
                if (pend) {
                    if (++spins > (1u << 24)) { *a.err = 1; break; }
                    if (spins > 64) __nanosleep(100);
                }
            }
#pragma unroll
            for (int t = 0; t < LCH; t++)
                if (want >> t & 1u) {
                    const int j = fwd ? (jBeg + c0 + t) : (jEnd - 1 - c0 - t);
                    if (fwd) {
#pragma unroll
                        for (int k = 0; k < 5; k++) dl[t][k] = sc[t] * dl[t][k];
                    }
                    const int js = j - stageLo;
                    if (js >= 0 && js < TMA_SE) lusgsSubSmem(xr, stage + (size_t)js * 25 * 32 + lane, dl[t]);
                    else subBlockGlobal(xr, a.offd + ((base + j) * 25) * 32 + lane, dl[t]);
                }
        }
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const double v = fwd ? xr[k] : rd * xr[k];
            __stcg((fwd ? a.y : a.z) + k * a.NPH + p, v);
            if (!fwd) a.x[k * a.NPH + p] = v;
        }
        __syncwarp();
        if (lane == 0) {
            __stcg((fwd ? a.hintF : a.hintR) + s, a.epoch);
            consumed[st] = i / TMA_STAGES + 1;
            __threadfence_block();
            mbarArrive(empty + st);  // stage may be refilled
        }
    }
}

// ---------------- tile mode: blocked wavefront ----------------
// Positions are ordered tile by tile (setup.cu).  One CTA sweeps a whole tile: the rows of a tile are processed
// intra-tile level by level with __syncthreads in between, in-tile neighbours come from shared memory, out-of-tile
// neighbours (always in tiles of lower tile level) from the sentinel-validated global buffers; the tile publishes one
// hint when it is complete.  The dependency chain across SMs shrinks from the number of levels to the number of tile
// levels (e.g. 598 -> 75 for 200^3), so the sweep becomes bandwidth bound.
struct TileArgs {
    int nTiles;
    const int *tileStart, *tileFPtr, *tileFLev, *tileRPtr, *tileRLev, *tileRRows, *sliceTile;
    const int *sliceOff, *rowNLow, *rowNInt, *col;
    const double *offd, *rD;
    double *x, *y, *z;
    size_t NPH;
    int *hintF, *hintR;
    int epoch;
    int* err;
    int maxRows;
};

__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void pollTile(const double* buf, const int* hint, const int* sliceTile, int epoch, size_t NPH, int q, double* out, int* err)
{
    unsigned int spins = 0;
    const int* h = hint + sliceTile[q >> 5];
    while (true) {
        if (ldHint(h) == epoch) {
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 5; k++) { out[k] = ldPoll(buf + k * NPH + q); ok &= !isSentinel(out[k]); }
            if (ok) return;
        }
        if (++spins > (1u << 24)) { *err = 1; return; }
        if (spins > 64) __nanosleep(100);
    }
}

__device__ __forceinline__ void subBlockGlobal(double* xr, const double* __restrict__ blk, const double* dl)
{
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const double b0 = __ldcs(blk + (size_t)(r * 5 + 0) * 32), b1 = __ldcs(blk + (size_t)(r * 5 + 1) * 32), b2 = __ldcs(blk + (size_t)(r * 5 + 2) * 32),
                     b3 = __ldcs(blk + (size_t)(r * 5 + 3) * 32), b4 = __ldcs(blk + (size_t)(r * 5 + 4) * 32);
        xr[r] -= b0 * dl[0];
        xr[r] -= b4 * dl[4];
        xr[r] -= b1 * dl[1] + b2 * dl[2] + b3 * dl[3];
    }
}

constexpr int TILE_TPB = 128;

// one row's sweep inputs, fetched one intra-tile level ahead (everything here is independent of the sweep values, except
// `own` in the reverse sweep which is validated against the sentinel before use)
struct RowPf {
    int r, n;             // local row (-1: none), number of neighbour entries of this sweep
    int jFirst;           // first entry index (forward: 0; reverse: nInt-1, descending)
    int q[LCH];
    double sc[LCH];
    double B[LCH][25];
    double own[5], rd;
    size_t base;
    int ln;
};

constexpr int TILE_CW = 8;        // column entries per row staged in shared memory (hex rows have 6)
constexpr int TILE_MAXHALO = 768; // out-of-tile neighbour values staged per tile and sweep

// shared-memory view of one tile: sweep values + the small per-row metadata, staged once per tile with coalesced loads
// so that the per-level critical path contains no dependent global index loads
struct TileSmem {
    double* xs;   // [5][MR] sweep values of the tile's rows
    double* rDs;  // [MR]
    int* nLow;    // [MR]
    int* nInt;    // [MR]
    int* cols;    // [MR/32][TILE_CW][32]  in-tile: position; out-of-tile: -(halo slot + 1)
    double* halo; // [5][TILE_MAXHALO] out-of-tile neighbour values of this sweep (forward: already scaled by rD)
    int* haloQ;   // [TILE_MAXHALO]
    int* cnt;     // [1]
};

__device__ __forceinline__ TileSmem carve(double* base, int MR)
{
    TileSmem t;
    t.xs = base;
    t.rDs = base + 5 * (size_t)MR;
    t.nLow = (int*)(t.rDs + MR);
    t.nInt = t.nLow + MR;
    t.cols = t.nInt + MR;
    t.halo = (double*)(t.cols + (size_t)(MR / 32) * TILE_CW * 32 + 2);
    t.halo = (double*)(((size_t)t.halo + 7) & ~(size_t)7);
    t.haloQ = (int*)(t.halo + 5 * TILE_MAXHALO);
    t.cnt = t.haloQ + TILE_MAXHALO;
    return t;
}

template <bool FWD>
__device__ __forceinline__ void pfLoad(const TileArgs& a, const TileSmem& sm, int t0, int t1, int r, RowPf& d)
{
    d.r = r;
    if (r < 0) { d.n = 0; return; }
    const int p = t0 + r;
    d.ln = p & 31;
    d.base = (size_t)a.sliceOff[p >> 5];
    const int nLow = sm.nLow[r];
    if (FWD) { d.n = nLow; d.jFirst = 0; d.rd = 1.0; }
    else { const int nInt = sm.nInt[r]; d.n = nInt - nLow; d.jFirst = nInt - 1; d.rd = sm.rDs[r]; }
    const double* src = FWD ? a.x : a.y;
#pragma unroll
    for (int k = 0; k < 5; k++) d.own[k] = FWD ? src[k * a.NPH + p] : ldPoll(src + k * a.NPH + p);
#pragma unroll
    for (int t = 0; t < LCH; t++) {
        d.q[t] = -1;
        if (t < d.n) {
            const int j = FWD ? (d.jFirst + t) : (d.jFirst - t);
            const int q = (j < TILE_CW) ? sm.cols[((r >> 5) * TILE_CW + j) * 32 + d.ln] : a.col[(d.base + j) * 32 + d.ln];
            d.q[t] = q;
            d.sc[t] = FWD ? ((q >= t0) ? sm.rDs[q - t0] : (q >= 0 ? a.rD[q] : 1.0)) : 1.0;
            const double* blk = a.offd + ((d.base + j) * 25) * 32 + d.ln;
#pragma unroll
            for (int k = 0; k < 25; k++) d.B[t][k] = __ldcs(blk + (size_t)k * 32);
        }
    }
}

template <bool FWD>
__device__ __forceinline__ void pfProcess(const TileArgs& a, const TileSmem& sm, int t0, int t1, const RowPf& d, int MR)
{
    double* xs = sm.xs;
    if (d.r < 0) return;
    const int p = t0 + d.r;
    double xr[5];
#pragma unroll
    for (int k = 0; k < 5; k++) xr[k] = d.own[k];
    if (!FWD) {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 5; k++) ok &= !isSentinel(xr[k]);
        if (!ok) pollTile(a.y, a.hintF, a.sliceTile, a.epoch, a.NPH, p, xr, a.err);  // own forward value not published yet
    }
#pragma unroll
    for (int t = 0; t < LCH; t++) {
        if (t < d.n) {
            double dl[5];
            const int q = d.q[t];
            if (q < 0) {  // staged out-of-tile value (forward: rD_q x_q already applied)
#pragma unroll
                for (int k = 0; k < 5; k++) dl[k] = sm.halo[k * TILE_MAXHALO + (-q - 1)];
            } else {
                const bool inTile = FWD ? (q >= t0) : (q < t1);
                if (inTile) {
#pragma unroll
                    for (int k = 0; k < 5; k++) dl[k] = xs[k * MR + (q - t0)];
                } else {
                    pollTile(FWD ? a.y : a.z, FWD ? a.hintF : a.hintR, a.sliceTile, a.epoch, a.NPH, q, dl, a.err);
                }
                if (FWD) {
#pragma unroll
                    for (int k = 0; k < 5; k++) dl[k] = d.sc[t] * dl[k];  // dW*_q = rD_q x_q (lusgs.C:194-216)
                }
            }
            lusgsSubReg(xr, d.B[t], dl);
        }
    }
    for (int t = LCH; t < d.n; t++) {  // rows with more than LCH neighbours in this sweep (polyhedral cells)
        const int j = FWD ? (d.jFirst + t) : (d.jFirst - t);
        const int q = (j < TILE_CW) ? sm.cols[((d.r >> 5) * TILE_CW + j) * 32 + d.ln] : a.col[(d.base + j) * 32 + d.ln];
        double dl[5];
        if (q < 0) {
#pragma unroll
            for (int k = 0; k < 5; k++) dl[k] = sm.halo[k * TILE_MAXHALO + (-q - 1)];
        } else {
            const bool inTile = FWD ? (q >= t0) : (q < t1);
            if (inTile) {
#pragma unroll
                for (int k = 0; k < 5; k++) dl[k] = xs[k * MR + (q - t0)];
            } else {
                pollTile(FWD ? a.y : a.z, FWD ? a.hintF : a.hintR, a.sliceTile, a.epoch, a.NPH, q, dl, a.err);
            }
            if (FWD) {
                const double sc = a.rD[q];
#pragma unroll
                for (int k = 0; k < 5; k++) dl[k] = sc * dl[k];
            }
        }
        subBlockGlobal(xr, a.offd + ((d.base + j) * 25) * 32 + d.ln, dl);
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const double v = FWD ? xr[k] : d.rd * xr[k];
        xs[k * MR + d.r] = v;
        __stcg((FWD ? a.y : a.z) + k * a.NPH + p, v);
        if (!FWD) a.x[k * a.NPH + p] = v;
    }
}

// Two thread groups alternate over the intra-tile levels: while group g computes level L from registers, the other
// group's loads for level L+1 are in flight; after the barrier g issues its loads for level L+2.
template <bool FWD>
__device__ __forceinline__ void sweepTile(const TileArgs& a, int tile, double* smemBase, int MR)
{
    const TileSmem sm = carve(smemBase, MR);
    if (threadIdx.x == 0) *sm.cnt = 0;
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int GT = TILE_TPB / 2;  // threads per group
    const int grp = tid / GT, gtid = tid % GT;
    const int t0 = a.tileStart[tile], t1 = a.tileStart[tile + 1];
    // pull this sweep's blocks and columns of the tile's slices into L2 while the predecessors finish
    for (int s = (t0 >> 5) + warp; s < (t1 >> 5); s += TILE_TPB / 32) {
        int nl = a.rowNLow[s * 32 + lane], ni = a.rowNInt[s * 32 + lane];
        int lo, hi;
        if (FWD) { lo = 0; hi = nl; }
        else { lo = (ni > nl) ? nl : (1 << 20); hi = ni; }
        for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
        if (hi > lo) {
            const size_t e0 = (size_t)a.sliceOff[s] + lo;
            const char* b = (const char*)(a.offd + e0 * 25 * 32);
            const size_t bytes = (size_t)(hi - lo) * 25 * 32 * 8;
            for (size_t off = (size_t)lane * 128; off < bytes; off += 32 * 128) prefetchL2(b + off);
        }
        // stage this slice's columns (first TILE_CW entries) in shared memory
        const int wdt = a.sliceOff[s + 1] - a.sliceOff[s];
        const int sl = s - (t0 >> 5);
        const int myLow = a.rowNLow[s * 32 + lane], myInt = a.rowNInt[s * 32 + lane];
        for (int j = 0; j < min(wdt, TILE_CW); j++) {
            int q = a.col[((size_t)a.sliceOff[s] + j) * 32 + lane];
            // entries of THIS sweep that point outside the tile get a halo slot
            const bool mineSweep = FWD ? (j < myLow) : (j >= myLow && j < myInt);
            if (mineSweep && (FWD ? (q < t0) : (q >= t1))) {
                const int slot = atomicAdd(sm.cnt, 1);
                if (slot < TILE_MAXHALO) { sm.haloQ[slot] = q; q = -(slot + 1); }
            }
            sm.cols[(sl * TILE_CW + j) * 32 + lane] = q;
        }
    }
    for (int r = tid; r < t1 - t0; r += TILE_TPB) { sm.nLow[r] = a.rowNLow[t0 + r]; sm.nInt[r] = a.rowNInt[t0 + r]; sm.rDs[r] = a.rD[t0 + r]; }
    __syncthreads();
    {
        // wait for the predecessor tiles once and stage every out-of-tile neighbour value of this sweep
        const int nH = min(*sm.cnt, TILE_MAXHALO);
        for (int h = tid; h < nH; h += TILE_TPB) {
            const int q = sm.haloQ[h];
            double v[5];
            pollTile(FWD ? a.y : a.z, FWD ? a.hintF : a.hintR, a.sliceTile, a.epoch, a.NPH, q, v, a.err);
            const double sc = FWD ? a.rD[q] : 1.0;
#pragma unroll
            for (int k = 0; k < 5; k++) sm.halo[k * TILE_MAXHALO + h] = FWD ? sc * v[k] : v[k];
        }
    }
    __syncthreads();
    const int* lev = FWD ? a.tileFLev : a.tileRLev;
    const int lp0 = FWD ? a.tileFPtr[tile] : a.tileRPtr[tile];
    const int nLev = (FWD ? a.tileFPtr[tile + 1] : a.tileRPtr[tile + 1]) - lp0 - 1;
    auto rowOf = [&](int L, int i) -> int {  // i-th row of level L handled by this thread, or -1
        if (L >= nLev) return -1;
        const int r0 = lev[lp0 + L], r1 = lev[lp0 + L + 1];
        const int idx = r0 + i * GT + gtid;
        if (idx >= r1) return -1;
        return FWD ? idx : a.tileRRows[t0 + idx];
    };
    RowPf cur;
    pfLoad<FWD>(a, sm, t0, t1, rowOf(grp, 0), cur);  // group 0 starts with level 0, group 1 with level 1
    for (int L = 0; L < nLev; L++) {
        const bool mine = (L & 1) == grp;
        if (mine) {
            pfProcess<FWD>(a, sm, t0, t1, cur, MR);
            const int width = lev[lp0 + L + 1] - lev[lp0 + L];
            for (int i = 1; i * GT < width; i++) {  // levels wider than a group: remaining rows, unpipelined
                pfLoad<FWD>(a, sm, t0, t1, rowOf(L, i), cur);
                pfProcess<FWD>(a, sm, t0, t1, cur, MR);
            }
        }
        __syncthreads();
        if (mine) pfLoad<FWD>(a, sm, t0, t1, rowOf(L + 2, 0), cur);
    }
    if (tid == 0) __stcg((FWD ? a.hintF : a.hintR) + tile, a.epoch);
    __syncthreads();  // the staged metadata is reused by the next tile
}

__global__ void __launch_bounds__(TILE_TPB)
k_lusgs_tile(TileArgs a)
{
    extern __shared__ double xs[];  // [5][maxRows]
    const int MR = a.maxRows;
    for (int tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) sweepTile<true>(a, tile, xs, MR);        // forward, ascending
    for (int tt = blockIdx.x; tt < a.nTiles; tt += gridDim.x) sweepTile<false>(a, a.nTiles - 1 - tt, xs, MR);  // reverse, descending
}


// ---------------- tile mode with TMA staging (k_lusgs_tile_tma): the chain-bound regime ----------------
// 64-row tiles (4x4x4 cells of a structured-like mesh), thread = row.  Each CTA runs TT_LANES independent "lanes" of 64
// consumer threads; a lane sweeps one tile at a time out of its own shared-memory stage.  One producer warp streams
// EVERYTHING a tile's sweep needs that does not depend on other tiles — the 5x5 blocks, column indices, row lengths, rD,
// the rows' intra-tile levels and their right-hand side / forward values — with cp.async.bulk into the stage of whichever
// lane has just released it, so the load of one lane overlaps the sweeps of the others.  A lane walks its tile's ~10
// intra-tile levels with a named barrier per level, reading only shared memory; out-of-tile neighbour values (always from
// tiles of a lower tile level) are polled once per tile with the hint + sentinel protocol of the level kernels.  The
// cross-SM dependency chain shrinks from the number of levels (3n) to the number of tile levels (3n/4), and an intra-tile
// level costs a barrier + ~100 shared-memory reads instead of an L2 round trip.  Meant for small meshes / partitions,
// where the level pipeline is bound by hop latency; bit-identical to it.
constexpr int TT_ROWS = 64;          // rows per tile = consumer threads per lane
constexpr int TT_SL = TT_ROWS / 32;  // slices per tile
constexpr int TT_SE = 3;             // staged block entries per slice (interior rows: exactly 3 per sweep; else global fallback)
constexpr int TT_CW = 8;             // staged column entries per slice
constexpr int TT_LANES = 4;          // concurrent tiles per CTA

struct TTStage {
    double blocks[TT_SL][TT_SE][25][32];  // 38400 B
    int cols[TT_SL][TT_CW][32];           //  2048 B
    int nLow[TT_ROWS], nInt[TT_ROWS], lev[TT_ROWS];
    double rD[TT_ROWS];
    double own[5][TT_ROWS];               // forward: right-hand side; reverse: the rows' forward values (validated)
    int desc[16];                         // the tile's descriptor (TT_DESC_*): consumers need no global metadata
    double xs[5][TT_ROWS];                // sweep values of the tile (not staged)
};
// descriptor layout (setup.cu): [0] t0 [1] nRows [2] nLev; per slice sl: [4+sl] first entry offset (sliceOff), [6+sl] first
// staged entry, [8+sl] staged entries, [10+sl] staged column entries
struct TTShared {
    TTStage st[TT_LANES];
    unsigned long long full[TT_LANES], empty[TT_LANES];
};

struct TileTmaArgs {
    int nTiles;
    const int *sliceTile, *rowLevF, *rowLevR, *tileDescF, *tileDescR;
    const int *rowNLow, *rowNInt, *col;
    const double *offd, *rD;
    double *x, *y, *z;
    size_t NPH;
    int *hintF, *hintR;
    int epoch;
    int* err;
};

__device__ __forceinline__ void laneBarrier(int ln) { asm volatile("bar.sync %0, %1;" ::"r"(ln + 1), "n"(TT_ROWS) : "memory"); }

__global__ void __launch_bounds__(TT_LANES * (TT_ROWS + 32), 1)
k_lusgs_tile_tma(TileTmaArgs a)
{
    extern __shared__ __align__(128) unsigned char ttRaw[];
    TTShared& sm = *reinterpret_cast<TTShared*>(ttRaw);
    const int tid = threadIdx.x, lane = tid & 31;
    const int G = gridDim.x, b = blockIdx.x, GL = G * TT_LANES;
    if (tid == 0) {
        for (int s2 = 0; s2 < TT_LANES; s2++) { mbarInit(sm.full + s2, 1); mbarInit(sm.empty + s2, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // lane ln of CTA b sweeps the tiles T = b + ln*G + i*GL (forward, ascending) and then nTiles-1-T (reverse): consecutive
    // tiles — same tile level — go to different SMs first
    auto nMineOf = [&](int ln) { const int first = b + ln * G; return (a.nTiles > first) ? (a.nTiles - first + GL - 1) / GL : 0; };
    auto itemTile = [&](int ln, int nMine, int i, bool& fwd) {
        fwd = i < nMine;
        const int T = b + ln * G + (fwd ? i : i - nMine) * GL;
        return fwd ? T : a.nTiles - 1 - T;
    };

    if (tid >= TT_LANES * TT_ROWS) {
        // ---------------- producer warps: warp w serves lane w; the next tile's descriptor is fetched while the current
        // one is being swept, so a released stage is refilled without a global round trip ----------------
        const int ln = (tid - TT_LANES * TT_ROWS) >> 5;
        if (lane == 0) {
            const int nMine = nMineOf(ln), nItems = 2 * nMine;
            TTStage& S = sm.st[ln];
            int4 d0, d1, d2;
            auto loadDesc = [&](int i) {
                bool fwd;
                const int tile = itemTile(ln, nMine, i, fwd);
                const int4* dp = reinterpret_cast<const int4*>((fwd ? a.tileDescF : a.tileDescR) + (size_t)16 * tile);
                d0 = dp[0]; d1 = dp[1]; d2 = dp[2];
            };
            if (nItems > 0) loadDesc(0);
            for (int i = 0; i < nItems; i++) {
                if (i > 0 && !mbarWait(sm.empty + ln, (i + 1) & 1, a.err)) break;  // stage still in use by item i-1
                bool fwd;
                const int tile = itemTile(ln, nMine, i, fwd);
                const int t0 = d0.x, nRows = d0.y, nSl = nRows >> 5;
                const int e0[2] = {d1.x, d1.y}, lo[2] = {d1.z, d1.w}, ne[2] = {d2.x, d2.y}, nc[2] = {d2.z, d2.w};
                unsigned bytes = (unsigned)nRows * (4 + 4 + 4 + 8 + 5 * 8) + 64;
                for (int sl = 0; sl < nSl; sl++) bytes += (unsigned)ne[sl] * 25 * 32 * 8 + (unsigned)nc[sl] * 32 * 4;
                mbarExpectTx(sm.full + ln, bytes);
                tmaLoad1D(S.desc, (fwd ? a.tileDescF : a.tileDescR) + (size_t)16 * tile, 64, sm.full + ln);
                for (int sl = 0; sl < nSl; sl++) {
                    if (ne[sl] > 0) tmaLoad1D(&S.blocks[sl][0][0][0], a.offd + ((size_t)e0[sl] + lo[sl]) * 25 * 32, (unsigned)ne[sl] * 25 * 32 * 8, sm.full + ln);
                    if (nc[sl] > 0) tmaLoad1D(&S.cols[sl][0][0], a.col + (size_t)e0[sl] * 32, (unsigned)nc[sl] * 32 * 4, sm.full + ln);
                }
                tmaLoad1D(S.nLow, a.rowNLow + t0, (unsigned)nRows * 4, sm.full + ln);
                tmaLoad1D(S.nInt, a.rowNInt + t0, (unsigned)nRows * 4, sm.full + ln);
                tmaLoad1D(S.lev, (fwd ? a.rowLevF : a.rowLevR) + t0, (unsigned)nRows * 4, sm.full + ln);
                tmaLoad1D(S.rD, a.rD + t0, (unsigned)nRows * 8, sm.full + ln);
                const double* src = fwd ? a.x : a.y;
                for (int k = 0; k < 5; k++) tmaLoad1D(&S.own[k][0], src + k * a.NPH + t0, (unsigned)nRows * 8, sm.full + ln);
                if (i + 1 < nItems) loadDesc(i + 1);
            }
        }
        return;
    }
    // ---------------- consumers: lane ln, thread r owns row t0 + r of the lane's current tile ----------------
    const int ln = tid / TT_ROWS, r = tid % TT_ROWS;
    const int nMine = nMineOf(ln), nItems = 2 * nMine;
    TTStage& S = sm.st[ln];
    for (int i = 0; i < nItems; i++) {
        bool fwd;
        const int tile = itemTile(ln, nMine, i, fwd);
        const double* buf = fwd ? a.y : a.z;
        const int* hint = fwd ? a.hintF : a.hintR;
        const int sl = r >> 5;
        mbarWait(sm.full + ln, i & 1, a.err);
        const int t0 = S.desc[0], t1 = t0 + S.desc[1], nLev = S.desc[2];
        const int p = t0 + r;
        const int sle = min(sl, (S.desc[1] >> 5) - 1);
        const int stageLo = S.desc[6 + sle];
        const size_t sliceE0 = (size_t)S.desc[4 + sle];
        const bool inRange = r < t1 - t0;
        const int myLev = inRange ? S.lev[r] : -1;  // -1: padding row, never swept
        const bool active = myLev >= 0;
        const int nLow = active ? S.nLow[r] : 0, nInt = active ? S.nInt[r] : 0;
        const int jBeg = fwd ? 0 : nLow, jEnd = fwd ? nLow : nInt;
        const int n = min(jEnd - jBeg, LCH);
        const double rd = (active && !fwd) ? S.rD[r] : 1.0;
        double xr[5] = {0, 0, 0, 0, 0};
        if (active) {
#pragma unroll
            for (int k = 0; k < 5; k++) xr[k] = S.own[k][r];
        }
        int q[LCH], jj[LCH];
        bool inTile[LCH];
        double dlo[LCH][5];
#pragma unroll
        for (int t = 0; t < LCH; t++) {
            q[t] = -1; jj[t] = 0; inTile[t] = true;
            if (t < n) {
                jj[t] = fwd ? (jBeg + t) : (jEnd - 1 - t);
                q[t] = S.cols[sl][jj[t]][lane];
                inTile[t] = q[t] >= t0 && q[t] < t1;
            }
        }
        // out-of-tile neighbours: always in tiles of a lower tile level (forward) / higher (reverse).  All hints of the thread
        // are polled together, then all values together (two L2 round trips per tile when the predecessors are done)
        double sco[LCH];
        unsigned int pend = 0;
#pragma unroll
        for (int t = 0; t < LCH; t++) {
            const bool out = t < n && !inTile[t];
            sco[t] = (out && fwd) ? a.rD[q[t]] : 1.0;
            if (out) pend |= 1u << t;
        }
        {
            unsigned int spins = 0;
            while (pend) {
                int hv[LCH];
#pragma unroll
                for (int t = 0; t < LCH; t++) hv[t] = (pend >> t & 1u) ? ldHint(hint + a.sliceTile[q[t] >> 5]) : 0;
                unsigned int ready = 0;
#pragma unroll
                for (int t = 0; t < LCH; t++)
                    if ((pend >> t & 1u) && hv[t] == a.epoch) {
                        ready |= 1u << t;
#pragma unroll
                        for (int k = 0; k < 5; k++) dlo[t][k] = ldPoll(buf + k * a.NPH + q[t]);
                    }
#pragma unroll
                for (int t = 0; t < LCH; t++)
                    if (ready >> t & 1u) {
                        bool ok = true;
#pragma unroll
                        for (int k = 0; k < 5; k++) ok &= !isSentinel(dlo[t][k]);
                        if (ok) pend &= ~(1u << t);
                    }
                if (pend) {
                    if (++spins > (1u << 24)) { *a.err = 1; break; }
                    if (spins > 64) __nanosleep(100);
                }
            }
        }
        if (fwd) {
#pragma unroll
            for (int t = 0; t < LCH; t++)
                if (t < n && !inTile[t]) {
#pragma unroll
                    for (int k = 0; k < 5; k++) dlo[t][k] = sco[t] * dlo[t][k];  // dW*_q = rD_q x_q (lusgs.C:194-216)
                }
        }
        if (active && !fwd) {  // own forward value: normally published long before the bulk copy read it
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 5; k++) ok &= !isSentinel(xr[k]);
            if (!ok) pollTile(a.y, a.hintF, a.sliceTile, a.epoch, a.NPH, p, xr, a.err);
        }
        // the tile's levels, one lane barrier each
        for (int L = 0; L < nLev; L++) {
            if (myLev == L) {
#pragma unroll
                for (int t = 0; t < LCH; t++)
                    if (t < n) {
                        double dl[5];
                        if (inTile[t]) {
                            const int qr = q[t] - t0;
                            const double sc = fwd ? S.rD[qr] : 1.0;
#pragma unroll
                            for (int k = 0; k < 5; k++) dl[k] = fwd ? sc * S.xs[k][qr] : S.xs[k][qr];
                        } else {
#pragma unroll
                            for (int k = 0; k < 5; k++) dl[k] = dlo[t][k];
                        }
                        const int js = jj[t] - stageLo;
                        if (js >= 0 && js < TT_SE) lusgsSubSmem(xr, &S.blocks[sl][js][0][lane], dl);
                        else subBlockGlobal(xr, a.offd + ((sliceE0 + jj[t]) * 25) * 32 + lane, dl);
                    }
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const double v = fwd ? xr[k] : rd * xr[k];
                    S.xs[k][r] = v;
                    __stcg((fwd ? a.y : a.z) + k * a.NPH + p, v);
                    if (!fwd) a.x[k * a.NPH + p] = v;
                }
            }
            laneBarrier(ln);
        }
        if (nLev == 0) laneBarrier(ln);
        if (r == 0) {
            __stcg((fwd ? a.hintF : a.hintR) + tile, a.epoch);
            mbarArrive(sm.empty + ln);  // every consumer of the lane is past its last read of the stage
        }
    }
}

__global__ void k_fill_sentinel(size_t n, unsigned long long* __restrict__ a, unsigned long long* __restrict__ b)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { a[i] = LUSGS_SENTINEL; b[i] = LUSGS_SENTINEL; }
}

// ------------------------------------------------------------------------------------------------ Jacobi
// LUscalarMatrix(J).inv() per cell, dense block in the reference's variable order (rho, rhoE, rhoU) — JacobiSmoother.C:42-101
__global__ void k_jacobi_invert(int NP, const int* __restrict__ pos2cell, const double* __restrict__ diag, double* __restrict__ invD)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    const int perm[5] = {0, 4, 1, 2, 3};  // reference index -> block index
    double a[5][5];
    for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) a[i][j] = diag[(size_t)(perm[i] * 5 + perm[j]) * NP + p];
    int piv[5];
    double vv[5];
    for (int i = 0; i < 5; i++) {
        double largest = 0.0;
        for (int j = 0; j < 5; j++) largest = fmax(largest, fabs(a[i][j]));
        vv[i] = 1.0 / largest;
    }
    for (int j = 0; j < 5; j++) {
        for (int i = 0; i < j; i++) { double sum = a[i][j]; for (int k = 0; k < i; k++) sum -= a[i][k] * a[k][j]; a[i][j] = sum; }
        int iMax = 0;
        double largest = 0.0;
        for (int i = j; i < 5; i++) {
            double sum = a[i][j];
            for (int k = 0; k < j; k++) sum -= a[i][k] * a[k][j];
            a[i][j] = sum;
            const double temp = vv[i] * fabs(sum);
            if (temp >= largest) { largest = temp; iMax = i; }
        }
        piv[j] = iMax;
        if (j != iMax) { for (int k = 0; k < 5; k++) { double t = a[iMax][k]; a[iMax][k] = a[j][k]; a[j][k] = t; } vv[iMax] = vv[j]; }
        if (a[j][j] == 0.0) a[j][j] = ICS_SMALL;
        if (j != 4) { const double rDiag = 1.0 / a[j][j]; for (int i = j + 1; i < 5; i++) a[i][j] *= rDiag; }
    }
    for (int cc = 0; cc < 5; cc++) {
        double x[5] = {0, 0, 0, 0, 0};
        x[cc] = 1.0;
        int ii = 0;
        for (int i = 0; i < 5; i++) {
            const int ip = piv[i];
            double sum = x[ip];
            x[ip] = x[i];
            if (ii != 0) { for (int j = ii - 1; j < i; j++) sum -= a[i][j] * x[j]; }
            else if (sum != 0.0) ii = i + 1;
            x[i] = sum;
        }
        for (int i = 4; i >= 0; i--) {
            double sum = x[i];
            for (int j = i + 1; j < 5; j++) sum -= a[i][j] * x[j];
            x[i] = sum / a[i][i];
        }
        for (int i = 0; i < 5; i++) invD[(size_t)(i * 5 + cc) * NP + p] = x[i];  // stored in reference order
    }
}

// x = D^-1 b  (JacobiSmoother::smooth with zero initial guess; JacobiSmoother.C:123-203)
__global__ void k_jacobi_apply(int NP, const int* __restrict__ pos2cell, const double* __restrict__ invD, double* __restrict__ x, size_t NPH)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP || pos2cell[p] < 0) return;
    const int perm[5] = {0, 4, 1, 2, 3};
    double var[5], res[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++) var[i] = -(0.0 - x[perm[i] * NPH + p]);
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) res[i] += invD[(size_t)(i * 5 + j) * NP + p] * var[j];
    for (int i = 0; i < 5; i++) x[perm[i] * NPH + p] = res[i];
}

// ------------------------------------------------------------------------------------------------ reductions
constexpr int RED_BLOCK = 256;

__device__ __forceinline__ double warpSum(double v)
{
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level tree, then the last block to finish combines the per-block partials in index order (deterministic)
template <int NV>
__device__ __forceinline__ void finishReduction(double* v, double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ out)
{
    __shared__ double sh[NV][RED_BLOCK / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // blockIdx.y = time instance (Harmonic Balance): separate partials, counter and outputs per instance
    partial += (size_t)blockIdx.y * NV * gridDim.x;
    counter += blockIdx.y;
    out += (size_t)blockIdx.y * NV;
#pragma unroll
    for (int k = 0; k < NV; k++) { double s = warpSum(v[k]); if (lane == 0) sh[k][warp] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int w2 = 0; w2 < RED_BLOCK / 32; w2++) s += sh[k][w2];
            partial[(size_t)k * gridDim.x + blockIdx.x] = s;
        }
        __threadfence();
        const unsigned int done = atomicAdd(counter, 1u);
        last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        for (int k = 0; k < NV; k++) {
            double s = 0.0;
            for (int i = threadIdx.x; i < gridDim.x; i += RED_BLOCK) s += __ldcg(partial + (size_t)k * gridDim.x + i);
            s = warpSum(s);
            __syncthreads();
            if (lane == 0) sh[k][warp] = s;
            __syncthreads();
            if (threadIdx.x == 0) {
                double t = 0.0;
                for (int w2 = 0; w2 < RED_BLOCK / 32; w2++) t += sh[k][w2];
                out[k] = t;
            }
        }
        if (threadIdx.x == 0) *counter = 0;
    }
}

// optional fused MGS update  w -= h*vs  (h read from the device scalar area), then  out = sum_k w.vd  (vd == null: w.w)
__global__ void __launch_bounds__(RED_BLOCK)
k_axpy_dot(int NP, size_t NPH, double* __restrict__ w, const double* __restrict__ vs, const double* __restrict__ hPtr, const double* __restrict__ vd,
           double* __restrict__ partial, unsigned int* __restrict__ counter, double* __restrict__ out)
{
    double acc[1] = {0.0};
    const double h = vs ? *hPtr : 0.0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < NP; p += gridDim.x * blockDim.x) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            double wk = w[k * NPH + p];
            if (vs) { wk -= h * vs[k * NPH + p]; w[k * NPH + p] = wk; }
            const double o = vd ? vd[k * NPH + p] : wk;
            s += wk * o;
        }
        acc[0] += s;
    }
    finishReduction<1>(acc, partial, counter, out);
}

// 5 component sums of |r| (gSumMag / gSumCmptMag, gmres.C:1081-1098)
// inst != null (Harmonic Balance): grid.y = nO, the y-slice only sums the rows of its own time instance
__global__ void __launch_bounds__(RED_BLOCK)
k_sum_mag5(int NP, size_t NPH, const double* __restrict__ r, const int* __restrict__ inst, double* __restrict__ partial, unsigned int* __restrict__ counter,
           double* __restrict__ out)
{
    double acc[5] = {0, 0, 0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < NP; p += gridDim.x * blockDim.x) {
        if (inst && inst[p] != (int)blockIdx.y) continue;
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k] += fabs(r[k * NPH + p]);
    }
    finishReduction<5>(acc, partial, counter, out);
}

// plain sums of the 5 components of W (gAverage, gmres.C:812-823)
__global__ void __launch_bounds__(RED_BLOCK)
k_sum5(int NP, size_t stride, const double* __restrict__ r, const int* __restrict__ inst, double* __restrict__ partial, unsigned int* __restrict__ counter,
       double* __restrict__ out)
{
    double acc[5] = {0, 0, 0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < NP; p += gridDim.x * blockDim.x) {
        if (inst && inst[p] != (int)blockIdx.y) continue;
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k] += r[k * stride + p];
    }
    finishReduction<5>(acc, partial, counter, out);
}

// normalisation factors (gmres.C:839-851): sum(|Ax| + |b|) for rho, rhoE; sum(mag(Ax) + mag(b)) for rhoU
__global__ void __launch_bounds__(RED_BLOCK)
k_norm_factors(int NP, size_t NPH, const double* __restrict__ ax, const double* __restrict__ b, const int* __restrict__ inst, double* __restrict__ partial,
               unsigned int* __restrict__ counter, double* __restrict__ out)
{
    double acc[3] = {0, 0, 0};
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < NP; p += gridDim.x * blockDim.x) {
        if (inst && inst[p] != (int)blockIdx.y) continue;
        acc[0] += fabs(ax[p]) + fabs(b[p]);
        acc[1] += fabs(ax[4 * NPH + p]) + fabs(b[4 * NPH + p]);
        const double m1 = sqrt(ax[NPH + p] * ax[NPH + p] + ax[2 * NPH + p] * ax[2 * NPH + p] + ax[3 * NPH + p] * ax[3 * NPH + p]);
        const double m2 = sqrt(b[NPH + p] * b[NPH + p] + b[2 * NPH + p] * b[2 * NPH + p] + b[3 * NPH + p] * b[3 * NPH + p]);
        acc[2] += m1 + m2;
    }
    finishReduction<3>(acc, partial, counter, out);
}

// ------------------------------------------------------------------------------------------------ vector ops
__global__ void k_sub_avg(int NP, const int* __restrict__ pos2cell, size_t NPH, const double* __restrict__ W, size_t wstride, const double* __restrict__ sums,
                          const int* __restrict__ inst, double nTot, double* __restrict__ x)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    const bool valid = pos2cell[p] >= 0;
    const double* sm = sums + ((inst && valid) ? 5 * inst[p] : 0);
#pragma unroll
    for (int k = 0; k < 5; k++) x[k * NPH + p] = valid ? (W[k * wstride + p] - sm[k] / nTot) : 0.0;
}

__global__ void k_scale_div(int NP, size_t NPH, const double* __restrict__ w, const double* __restrict__ beta2, double* __restrict__ v)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    const double beta = sqrt(*beta2);
#pragma unroll
    for (int k = 0; k < 5; k++) v[k * NPH + p] = w[k * NPH + p] / beta;
}

__global__ void k_copy(size_t n, const double* __restrict__ a, double* __restrict__ b)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = a[i];
}
__global__ void k_set(size_t n, double v, double* __restrict__ b)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = v;
}

// dW += sum_i y_i v_i, one vector after the other (gmres.C:1044-1060)
__global__ void k_update_x(int NP, size_t NPH, int m, const double* __restrict__ kry, const double* __restrict__ yh, double* __restrict__ x)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        double v = x[k * NPH + p];
        for (int i = 0; i < m; i++) v += yh[i] * kry[(size_t)i * 5 * NPH + k * NPH + p];
        x[k * NPH + p] = v;
    }
}

__global__ void k_zero_dir(int NP, size_t NPH, int comp, double* __restrict__ x)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < NP) x[(size_t)(1 + comp) * NPH + p] = 0.0;
}

__global__ void k_co_update(int NP, double ratio, double coMin, double coMax, double* __restrict__ co)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    double v = co[p] * ratio;
    co[p] = fmax(fmin(v, coMax), coMin);
}

// ---- GMRES scalars (single thread; gmres.C:44-69, 1011-1041) ----
// layout of the scalar area (doubles): [0] beta^2 (norm of w)  [1..] h column (m)  [S_H] H (m*m)  [S_BH] bh (m+1)
// [S_C] c (m)  [S_S] s (m)  [S_Y] yh (m)
struct ScalLayout { int m, H, BH, C, S, Y, HCOL, BETA2; };
__host__ __device__ inline ScalLayout scalLayout(int m)
{
    ScalLayout L;
    L.m = m; L.BETA2 = 0; L.HCOL = 8; L.H = L.HCOL + m; L.BH = L.H + m * m; L.C = L.BH + m + 1; L.S = L.C + m; L.Y = L.S + m;
    return L;
}

__global__ void k_init_bh(ScalLayout L, double* __restrict__ sc)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int i = 0; i <= L.m; i++) sc[L.BH + i] = 0.0;
        sc[L.BH] = sqrt(sc[L.BETA2]);
    }
}

__global__ void k_givens(ScalLayout L, int i, double* __restrict__ sc)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int m = L.m;
    double* H = sc + L.H;
    double* c = sc + L.C;
    double* s = sc + L.S;
    double* bh = sc + L.BH;
    for (int j = 0; j <= i; j++) H[j * m + i] = sc[L.HCOL + j];
    const double beta = sqrt(sc[L.BETA2]);
    for (int j = 0; j < i; j++) {
        const double Hji = H[j * m + i];
        H[j * m + i] = c[j] * Hji - s[j] * H[(j + 1) * m + i];
        H[(j + 1) * m + i] = s[j] * Hji + c[j] * H[(j + 1) * m + i];
    }
    // givensRotation(H[i][i], beta, c[i], s[i])
    {
        const double h = H[i * m + i];
        if (beta == 0) { c[i] = 1; s[i] = 0; }
        else if (fabs(beta) > fabs(h)) { const double tau = -h / beta; s[i] = 1.0 / sqrt(1.0 + tau * tau); c[i] = s[i] * tau; }
        else { const double tau = -beta / h; c[i] = 1.0 / sqrt(1.0 + tau * tau); s[i] = c[i] * tau; }
    }
    const double bhi = bh[i];
    bh[i] = c[i] * bhi - s[i] * bh[i + 1];
    bh[i + 1] = s[i] * bhi + c[i] * bh[i + 1];
    H[i * m + i] = c[i] * H[i * m + i] - s[i] * beta;
}

__global__ void k_backsub(ScalLayout L, double* __restrict__ sc)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int m = L.m;
    const double* H = sc + L.H;
    const double* bh = sc + L.BH;
    double* yh = sc + L.Y;
    for (int i = m - 1; i >= 0; i--) {
        double sum = bh[i];
        for (int j = i + 1; j < m; j++) sum -= H[i * m + j] * yh[j];
        const double d = H[i * m + i];
        yh[i] = sum / (d < 0 ? d - ICS_VSMALL : d + ICS_VSMALL);  // stabilise(H[i][i], VSMALL)
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host drivers
int ics_spmv(icsb200_ctx* c, const double* x, double* y, const double* b)
{
    int r = ics_halo_fields(c, const_cast<double*>(x), c->NPH, 5, 1u << 1);   // components 1..3 = rhoU
    if (r) return r;
    LaunchScope ls(c, TM_SPMV);
    const int grid = gridFor(c->NP, 128);
    const HBSpmv hb = ics_hb_spmv_args(c);
    if (c->hbNO > 1) {
        if (b) k_spmv<true, true><<<grid, 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_offd, c->d_diag, x, c->NPH, b, y, hb);
        else k_spmv<false, true><<<grid, 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_offd, c->d_diag, x, c->NPH, nullptr, y, hb);
    } else {
        if (b) k_spmv<true, false><<<grid, 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_offd, c->d_diag, x, c->NPH, b, y, hb);
        else k_spmv<false, false><<<grid, 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_sliceOff, c->d_rowNAll, c->d_col, c->d_offd, c->d_diag, x, c->NPH, nullptr, y, hb);
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_lusgs(icsb200_ctx* c, double* x)
{
    if (!c->rDValid) { int r = ics_rdiag(c); if (r) return r; }
    if (c->blkMode) return ics_lusgs_blk(c, x);  // block tiles, in place (lusgs_blk.cu)
    const size_t V5 = (size_t)5 * c->NPH;
    if (c->lusgsGrid == 0) {
        int perSM = 0;
        CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_lusgs, 256, 0));
        if (perSM < 1) return ics_fail(c, ICSB200_ECUDA, "lusgs kernel does not fit on an SM");
        c->lusgsGrid = c->numSMs * perSM;
        int r = devAlloc(c, &c->d_lusgsYZ, 2 * V5);
        if (r) return r;
        if ((r = devAlloc(c, &c->d_lusgsHint, (size_t)2 * c->nSlices))) return r;
        CUDA_TRY(c, cudaMemsetAsync(c->d_lusgsHint, 0, sizeof(int) * 2 * c->nSlices, c->stream));
        c->lusgsEpoch = 0;
    }
    if (c->tileMode && c->tileTma) {
        TileTmaArgs t{};
        t.nTiles = c->nTiles;
        t.sliceTile = c->d_sliceTile;
        t.rowLevF = c->d_rowLevF; t.rowLevR = c->d_rowLevR; t.tileDescF = c->d_tileDescF; t.tileDescR = c->d_tileDescR;
        t.rowNLow = c->d_rowNLow; t.rowNInt = c->d_rowNInt; t.col = c->d_col;
        t.offd = c->d_offd; t.rD = c->d_rD; t.x = x; t.NPH = c->NPH;
        t.y = c->d_lusgsYZ; t.z = c->d_lusgsYZ + V5;
        t.hintF = c->d_lusgsHint; t.hintR = c->d_lusgsHint + c->nSlices; t.epoch = ++c->lusgsEpoch;
        t.err = (int*)c->d_counter + 48;
        const size_t smem = sizeof(TTShared) + 128;
        static bool attrSet = false;
        if (!attrSet) {
            CUDA_TRY(c, cudaFuncSetAttribute(k_lusgs_tile_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            attrSet = true;
        }
        const int grid = std::min(c->numSMs, std::max(1, (c->nTiles + TT_LANES - 1) / TT_LANES));
        LaunchScope ls(c, TM_LUSGS);
        k_fill_sentinel<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, (unsigned long long*)t.y, (unsigned long long*)t.z);
        c->launches++;
        void* targs[] = {&t};
        CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)k_lusgs_tile_tma, dim3(grid), dim3(TT_LANES * (TT_ROWS + 32)), targs, smem, c->stream));
        return 0;
    }
    if (c->tileMode) {
        TileArgs t{};
        t.nTiles = c->nTiles;
        t.tileStart = c->d_tileStart; t.tileFPtr = c->d_tileFPtr; t.tileFLev = c->d_tileFLev; t.tileRPtr = c->d_tileRPtr;
        t.tileRLev = c->d_tileRLev; t.tileRRows = c->d_tileRRows; t.sliceTile = c->d_sliceTile;
        t.sliceOff = c->d_sliceOff; t.rowNLow = c->d_rowNLow; t.rowNInt = c->d_rowNInt; t.col = c->d_col;
        t.offd = c->d_offd; t.rD = c->d_rD; t.x = x; t.NPH = c->NPH;
        t.y = c->d_lusgsYZ; t.z = c->d_lusgsYZ + V5;
        t.hintF = c->d_lusgsHint; t.hintR = c->d_lusgsHint + c->nSlices; t.epoch = ++c->lusgsEpoch;
        t.err = (int*)c->d_counter + 48;
        t.maxRows = c->tileMaxRows;
        const size_t smem = (size_t)c->tileMaxRows * (6 * sizeof(double) + (2 + TILE_CW) * sizeof(int)) + 5 * TILE_MAXHALO * sizeof(double) + (TILE_MAXHALO + 8) * sizeof(int) + 64;
        if (c->lusgsTileGrid == 0) {
            int perSM = 0;
            CUDA_TRY(c, cudaFuncSetAttribute(k_lusgs_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_lusgs_tile, TILE_TPB, smem));
            if (perSM < 1) return ics_fail(c, ICSB200_ECUDA, "lusgs tile kernel does not fit on an SM");
            c->lusgsTileGrid = c->numSMs * perSM;
        }
        int grid = std::min(c->lusgsTileGrid, c->nTiles);
        {
            static const char* e3 = getenv("ICSB200_LUSGS_GRID");
            if (e3) grid = std::min(grid, std::max(1, atoi(e3)));
        }
        LaunchScope ls(c, TM_LUSGS);
        k_fill_sentinel<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, (unsigned long long*)t.y, (unsigned long long*)t.z);
        c->launches++;
        void* targs[] = {&t};
        CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)k_lusgs_tile, dim3(grid), dim3(TILE_TPB), targs, smem, c->stream));
        return 0;
    }
    {
        static const char* impl = getenv("ICSB200_LUSGS_IMPL");
        if (!(impl && std::string(impl) == "reg")) {
            TmaArgs t{};
            t.nSlices = c->nSlices;
            t.sliceOff = c->d_sliceOff; t.rowNLow = c->d_rowNLow; t.rowNInt = c->d_rowNInt; t.col = c->d_col;
            t.sliceFwdHi = c->d_sliceRange; t.sliceRevLo = c->d_sliceRange + c->nSlices; t.sliceRevHi = c->d_sliceRange + 2 * (size_t)c->nSlices;
            t.offd = c->d_offd; t.rD = c->d_rD; t.x = x; t.NPH = c->NPH;
            t.y = c->d_lusgsYZ; t.z = c->d_lusgsYZ + V5;
            t.hintF = c->d_lusgsHint; t.hintR = c->d_lusgsHint + c->nSlices; t.epoch = ++c->lusgsEpoch;
            t.err = (int*)c->d_counter + 48;
            const size_t smem = (size_t)TMA_STAGES * TMA_STAGE_DOUBLES * sizeof(double) + 2 * TMA_STAGES * sizeof(unsigned long long) + TMA_STAGES * sizeof(int) + 128;
            if (!c->lusgsTmaReady) {
                CUDA_TRY(c, cudaFuncSetAttribute(k_lusgs_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                c->lusgsTmaReady = true;
            }
            int grid = std::min(c->numSMs, std::max(1, c->nSlices));
            LaunchScope ls(c, TM_LUSGS);
            k_fill_sentinel<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, (unsigned long long*)t.y, (unsigned long long*)t.z);
            c->launches++;
            void* targs[] = {&t};
            CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)k_lusgs_tma, dim3(grid), dim3((TMA_NC + 1) * 32), targs, smem, c->stream));
            return 0;
        }
    }
    LusgsArgs a{};
    a.nSlices = c->nSlices;
    a.sliceOff = c->d_sliceOff; a.rowNLow = c->d_rowNLow; a.rowNInt = c->d_rowNInt; a.col = c->d_col;
    a.offd = c->d_offd; a.rD = c->d_rD; a.x = x; a.NPH = c->NPH;
    a.y = c->d_lusgsYZ; a.z = c->d_lusgsYZ + V5;
    a.err = (int*)c->d_counter + 48;
    a.hintF = c->d_lusgsHint; a.hintR = c->d_lusgsHint + c->nSlices; a.epoch = ++c->lusgsEpoch;
    {
        static const char* e1 = getenv("ICSB200_LUSGS_SPIN");
        static const char* e2 = getenv("ICSB200_LUSGS_SLEEP");
        a.busySpins = e1 ? (unsigned)atoi(e1) : 64u;
        a.sleepNs = e2 ? (unsigned)atoi(e2) : 100u;
    }
    int grid = std::min(c->lusgsGrid, std::max(1, (c->nSlices + 7) / 8));
    {
        static const char* e3 = getenv("ICSB200_LUSGS_GRID");
        if (e3) grid = std::min(grid, std::max(1, atoi(e3)));
    }
    a.trace = nullptr;
    {
        static const char* e4 = getenv("ICSB200_LUSGS_TRACE");
        if (e4) {
            if (!c->d_lusgsTrace) { int r = devAlloc(c, &c->d_lusgsTrace, (size_t)4 * c->nSlices); if (r) return r; }
            cudaMemsetAsync(c->d_lusgsTrace, 0, sizeof(long long) * 4 * c->nSlices, c->stream);
            a.trace = c->d_lusgsTrace;
        }
    }
    LaunchScope ls(c, TM_LUSGS);
    k_fill_sentinel<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, (unsigned long long*)a.y, (unsigned long long*)a.z);
    c->launches++;
    void* args[] = {&a};
    CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)k_lusgs, dim3(grid), dim3(256), args, 0, c->stream));
    return 0;
}

// the sweeps flag a broken dependency schedule instead of hanging; checked once per solve
static int lusgsCheck(icsb200_ctx* c)
{
    int h = 0;
    CUDA_TRY(c, cudaMemcpyAsync(&h, (int*)c->d_counter + 48, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (h) {
        cudaMemsetAsync((int*)c->d_counter + 48, 0, sizeof(int), c->stream);
        return ics_fail(c, ICSB200_ECUDA, "lusgs: dependency wait timed out (non-finite values in the sweep?)");
    }
    return 0;
}

int ics_jacobi_prepare(icsb200_ctx* c)
{
    if (c->invDValid) return 0;
    if (!c->d_invD) { int r = devAlloc(c, &c->d_invD, (size_t)25 * c->NP); if (r) return r; }
    LaunchScope ls(c, TM_JACOBI);
    k_jacobi_invert<<<gridFor(c->NP, 128), 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_diag, c->d_invD);
    CUDA_TRY(c, cudaGetLastError());
    c->invDValid = true;
    return 0;
}

int ics_jacobi(icsb200_ctx* c, double* x)
{
    int r = ics_jacobi_prepare(c);
    if (r) return r;
    LaunchScope ls(c, TM_JACOBI);
    k_jacobi_apply<<<gridFor(c->NP, 128), 128, 0, c->stream>>>(c->NP, c->d_pos2cell, c->d_invD, x, c->NPH);
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

static int redGrid(const icsb200_ctx* c) { return std::min(4 * c->numSMs, std::max(1, gridFor(c->NP, RED_BLOCK))); }

// setCoAndDeltaT.H:3-37 — switched evolution relaxation of the pseudo Courant number
int ics_pseudo_ser(icsb200_ctx* c)
{
    if (c->hbNO > 1) {
        // dbnsFullyImplicitHBFoam/outerLoop.H:32-64 + setCoAndDeltaT.H:3-11.  As coded there, coNumRatio stays 0 on the
        // first iteration that has an initRes but no prevRes yet and the Courant field is still multiplied by it.
        if (!c->haveInitRes) return 0;
        const int nO = c->hbNO;
        double ratio = 0.0;
        if (!c->firstIter && c->havePrevRes) {
            auto sq = [](double x) { return x * x; };
            double ni = 0, np = 0;
            for (int J = 0; J < nO; J++) {
                const double* vi = &c->hbVInit[3 * J];
                const double* vp = &c->hbVInitPrev[3 * J];
                ni += sq(c->hbSInit[2 * J]) + sq(c->hbSInit[2 * J + 1]) + (vi[0] * vi[0] + vi[1] * vi[1] + vi[2] * vi[2]);
                np += sq(c->hbSInitPrev[2 * J]) + sq(c->hbSInitPrev[2 * J + 1]) + (vp[0] * vp[0] + vp[1] * vp[1] + vp[2] * vp[2]);
            }
            ratio = std::sqrt(np) / std::sqrt(ni);
            ratio = std::max(std::min(ratio, c->sch.pseudo_co_num_max_incr), c->sch.pseudo_co_num_min_decr);
        }
        c->hbSInitPrev = c->hbSInit;
        c->hbVInitPrev = c->hbVInit;
        c->havePrevRes = true;
        if (!c->firstIter) {
            LaunchScope ls(c, TM_UPDATE);
            k_co_update<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, ratio, c->sch.pseudo_co_num_min, c->sch.pseudo_co_num_max, c->d_co);
            CUDA_TRY(c, cudaGetLastError());
        }
        return 0;
    }
    if (c->haveInitRes) {
        if (!c->firstIter && c->havePrevRes) {
            const icsb200_residuals &ir = c->initRes, &pr = c->prevRes;
            auto sq = [](double x) { return x * x; };
            double normInit = std::sqrt(sq(ir.s_init[0]) + sq(ir.s_init[1]) + (ir.v_init[0] * ir.v_init[0] + ir.v_init[1] * ir.v_init[1] + ir.v_init[2] * ir.v_init[2]));
            double normPrev = std::sqrt(sq(pr.s_init[0]) + sq(pr.s_init[1]) + (pr.v_init[0] * pr.v_init[0] + pr.v_init[1] * pr.v_init[1] + pr.v_init[2] * pr.v_init[2]));
            double ratio = normPrev / normInit;
            ratio = std::max(std::min(ratio, c->sch.pseudo_co_num_max_incr), c->sch.pseudo_co_num_min_decr);
            if (c->sch.local_timestepping) {
                LaunchScope ls(c, TM_UPDATE);
                k_co_update<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->NP, ratio, c->sch.pseudo_co_num_min, c->sch.pseudo_co_num_max, c->d_co);
                CUDA_TRY(c, cudaGetLastError());
            } else {
                c->pseudoCoNum *= ratio;
                c->pseudoCoNum = std::max(std::min(c->pseudoCoNum, c->sch.pseudo_co_num_max), c->sch.pseudo_co_num_min);
            }
        }
        c->prevRes = c->initRes;
        c->havePrevRes = true;
    }
    return 0;
}

static int precond(icsb200_ctx* c, int kind, double* x)
{
    if (kind == ICSB200_PRECOND_LUSGS) return ics_lusgs(c, x);
    if (kind == ICSB200_PRECOND_JACOBI) return c->hbNO > 1 ? ics_hb_jacobi(c, x) : ics_jacobi(c, x);
    return ics_fail(c, ICSB200_EINVAL, "Unknown preconditioner; valid types are: LUSGS Jacobi");
}

// gmres::solveDelta (6-arg): W = d_Wprev, b = d_src, result dW = d_dW
int ics_gmres(icsb200_ctx* c, const icsb200_solver_controls* ctl, icsb200_residuals* res)
{
    const int m = ctl->n_directions;
    const bool smooth = ctl->solver == ICSB200_SOLVER_SMOOTH;   // smoothSolverCoupled: m = nSweeps, no Krylov space
    if (ctl->solver != ICSB200_SOLVER_GMRES && !smooth) return ics_fail(c, ICSB200_EINVAL, "Unknown solver; valid types are: GMRES smoothSolverCoupled");
    if (m < 1 || m > 50) return ics_fail(c, ICSB200_EINVAL, smooth ? "nSweeps out of range" : "nDirections out of range");
    if (smooth && c->hbNO > 1) return ics_fail(c, ICSB200_EINVAL, "smoothSolverCoupled is not available for the Harmonic Balance system");
    const int NP = c->NP;
    const size_t NPH = c->NPH, V5 = 5 * NPH;
    if (!smooth && c->mAlloc < m) {
        int r = devAlloc(c, &c->d_kry, (size_t)m * V5);
        if (r) return r;
        CUDA_TRY(c, cudaMemsetAsync(c->d_kry, 0, sizeof(double) * m * V5, c->stream));
        c->mAlloc = m;
    }
    std::memset(res, 0, sizeof(*res));
    const ScalLayout L = scalLayout(m);
    double* sc = c->d_scal;
    double* red = c->d_scal + 3000;  // reduction outputs that the host reads
    const int rg = redGrid(c);
    const int g256 = gridFor(NP, 256);
    int r;
    // Harmonic Balance: nI time instances, every per-variable quantity of the residualsIO exists once per instance
    const int nI = c->hbNO;
    const int* inst = nI > 1 ? c->d_hbInst : nullptr;
    const dim3 rgI(rg, nI);
    // layout of the host-read reduction area: [0,80) sums of W, [100] cell count, [128,176) norm factors,
    // [192,272) |b| sums, [288,368) |r| sums
    enum { R_SUM = 0, R_NTOT = 100, R_NORM = 128, R_B = 192, R_R = 288, R_END = 368 };
    // global cell count of one instance mesh
    double nTot = c->N / nI;
    if (c->nRanks > 1) {
        CUDA_TRY(c, cudaMemcpyAsync(red + R_NTOT, &nTot, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if ((r = allreduceSum(c, red + R_NTOT, 1))) return r;
        CUDA_TRY(c, cudaMemcpyAsync(&nTot, red + R_NTOT, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
    // ---- normalisation: A (W - avg(W))   (gmres.C:812-851)
    {
        LaunchScope ls(c, TM_RED);
        k_sum5<<<rgI, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_Wprev, inst, c->d_partial, c->d_counter, red + R_SUM);
    }
    if ((r = allreduceSum(c, red + R_SUM, 5 * nI))) return r;
    {
        LaunchScope ls(c, TM_VEC);
        k_sub_avg<<<g256, 256, 0, c->stream>>>(NP, c->d_pos2cell, NPH, c->d_Wprev, NPH, red + R_SUM, inst, nTot, c->d_x);
    }
    if ((r = ics_spmv(c, c->d_x, c->d_w, nullptr))) return r;
    {
        LaunchScope ls(c, TM_RED);
        k_norm_factors<<<rgI, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_w, c->d_src, inst, c->d_partial, c->d_counter, red + R_NORM);
        k_sum_mag5<<<rgI, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_src, inst, c->d_partial, c->d_counter, red + R_B);
        c->launches++;
    }
    if ((r = allreduceSum(c, red + R_NORM, 3 * nI))) return r;
    if ((r = allreduceSum(c, red + R_B, 5 * nI))) return r;
    CUDA_TRY(c, cudaMemcpyAsync(c->h_scal, red, sizeof(double) * R_END, cudaMemcpyDeviceToHost, c->stream));
    // x0 = 0, r0 = b
    {
        LaunchScope ls(c, TM_VEC);
        k_set<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, 0.0, c->d_x);
        k_copy<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, c->d_src, c->d_w);
        c->launches++;
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    // residualsIO per instance: scalars (rho_I, rhoE_I) and the vector rhoU_I
    std::vector<double> sNorm(2 * nI), vNorm(nI), sInit(2 * nI), vInit(3 * nI), sFinal(2 * nI), vFinal(3 * nI);
    for (int I = 0; I < nI; I++) {
        sNorm[2 * I] = c->h_scal[R_NORM + 3 * I] + ICS_VSMALL;
        sNorm[2 * I + 1] = c->h_scal[R_NORM + 3 * I + 1] + ICS_VSMALL;
        vNorm[I] = c->h_scal[R_NORM + 3 * I + 2] + ICS_VSMALL;
        sInit[2 * I] = c->h_scal[R_B + 5 * I] / sNorm[2 * I];
        sInit[2 * I + 1] = c->h_scal[R_B + 5 * I + 4] / sNorm[2 * I + 1];
        for (int d = 0; d < 3; d++) vInit[3 * I + d] = c->h_scal[R_B + 5 * I + 1 + d] / vNorm[I];
    }
    sFinal = sInit;
    vFinal = vInit;
    auto publish = [&]() {
        res->s_init[0] = sInit[0]; res->s_init[1] = sInit[1];
        res->s_final[0] = sFinal[0]; res->s_final[1] = sFinal[1];
        for (int d = 0; d < 3; d++) { res->v_init[d] = vInit[d]; res->v_final[d] = vFinal[d]; }
        if (nI > 1) { c->hbSInit = sInit; c->hbVInit = vInit; c->hbSFinal = sFinal; c->hbVFinal = vFinal; }
    };
    publish();
    bool stop = false;
    do {
      if (smooth) {
        // JacobiSmoother::smooth (JacobiSmoother.C:120-203): x <- D^-1 (b - (A - D) x), evaluated as x + D^-1 (b - A x) with the
        // full product and the block-Jacobi kernels (equal up to rounding; the solver bar is 1e-8)
        {
            LaunchScope ls(c, TM_VEC);
            k_set<<<1, 32, 0, c->stream>>>(1, 1.0, sc + L.Y);
        }
        for (int sweep = 0; sweep < m; sweep++) {
            if ((r = ics_spmv(c, c->d_x, c->d_w, c->d_src))) return r;
            if ((r = precond(c, ICSB200_PRECOND_JACOBI, c->d_w))) return r;
            LaunchScope ls(c, TM_VEC);
            k_update_x<<<g256, 256, 0, c->stream>>>(NP, NPH, 1, c->d_w, sc + L.Y, c->d_x);
        }
      } else {
        if ((r = precond(c, ctl->preconditioner, c->d_w))) return r;
        {
            LaunchScope ls(c, TM_RED);
            k_axpy_dot<<<rg, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_w, nullptr, nullptr, nullptr, c->d_partial, c->d_counter, sc + L.BETA2);
        }
        if ((r = allreduceSum(c, sc + L.BETA2, 1))) return r;
        {
            LaunchScope ls(c, TM_VEC);
            k_init_bh<<<1, 32, 0, c->stream>>>(L, sc);
        }
        for (int i = 0; i < m; i++) {
            double* vi = c->d_kry + (size_t)i * V5;
            {
                LaunchScope ls(c, TM_VEC);
                k_scale_div<<<g256, 256, 0, c->stream>>>(NP, NPH, c->d_w, sc + L.BETA2, vi);
            }
            if ((r = ics_spmv(c, vi, c->d_w, nullptr))) return r;
            if ((r = precond(c, ctl->preconditioner, c->d_w))) return r;
            // modified Gram-Schmidt: h_0 = w.v_0 ; then (w -= h_j v_j ; h_{j+1} = w.v_{j+1}) ... ; beta^2 = w.w
            {
                LaunchScope ls(c, TM_RED);
                k_axpy_dot<<<rg, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_w, nullptr, nullptr, c->d_kry, c->d_partial, c->d_counter, sc + L.HCOL);
            }
            if ((r = allreduceSum(c, sc + L.HCOL, 1))) return r;
            for (int j = 0; j <= i; j++) {
                const double* vj = c->d_kry + (size_t)j * V5;
                const double* vn = (j < i) ? c->d_kry + (size_t)(j + 1) * V5 : nullptr;
                double* out = (j < i) ? sc + L.HCOL + j + 1 : sc + L.BETA2;
                {
                    LaunchScope ls(c, TM_RED);
                    k_axpy_dot<<<rg, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_w, vj, sc + L.HCOL + j, vn, c->d_partial, c->d_counter, out);
                }
                if ((r = allreduceSum(c, out, 1))) return r;
            }
            {
                LaunchScope ls(c, TM_VEC);
                k_givens<<<1, 32, 0, c->stream>>>(L, i, sc);
            }
        }
        {
            LaunchScope ls(c, TM_VEC);
            k_backsub<<<1, 32, 0, c->stream>>>(L, sc);
            k_update_x<<<g256, 256, 0, c->stream>>>(NP, NPH, m, c->d_kry, sc + L.Y, c->d_x);
            c->launches++;
        }
      }
        // true residual r = b - A dW
        if ((r = ics_spmv(c, c->d_x, c->d_w, c->d_src))) return r;
        {
            LaunchScope ls(c, TM_RED);
            k_sum_mag5<<<rgI, RED_BLOCK, 0, c->stream>>>(NP, NPH, c->d_w, inst, c->d_partial, c->d_counter, red + R_R);
        }
        if ((r = allreduceSum(c, red + R_R, 5 * nI))) return r;
        CUDA_TRY(c, cudaMemcpyAsync(c->h_scal + R_R, red + R_R, sizeof(double) * 5 * nI, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        for (int I = 0; I < nI; I++) {
            sFinal[2 * I] = c->h_scal[R_R + 5 * I] / sNorm[2 * I];
            sFinal[2 * I + 1] = c->h_scal[R_R + 5 * I + 4] / sNorm[2 * I + 1];
            for (int d = 0; d < 3; d++) {
                vFinal[3 * I + d] = c->h_scal[R_R + 5 * I + 1 + d] / vNorm[I];
                if (c->solutionD[d] == -1) vFinal[3 * I + d] = 0.0;
            }
        }
        res->n_iterations += smooth ? m : 1;      // smoothSolverCoupled.C:511 counts sweeps, gmres.C:1104 restarts
        publish();
        // solver::stop (coupledMatrixSolver.C:198-221) with residualsIO::max / maxRel over every variable
        if (res->n_iterations < ctl->min_iter) stop = false;
        else {
            double mx = -ICS_VGREAT, mr = -ICS_VGREAT;
            for (int i = 0; i < 2 * nI; i++) { mx = std::max(mx, sFinal[i]); mr = std::max(mr, sFinal[i] / (sInit[i] + ICS_ROOTVSMALL)); }
            for (int I = 0; I < nI; I++) {
                mx = std::max(mx, std::max(vFinal[3 * I], std::max(vFinal[3 * I + 1], vFinal[3 * I + 2])));
                for (int d = 0; d < 3; d++) if (c->solutionD[d] == 1) mr = std::max(mr, vFinal[3 * I + d] / (vInit[3 * I + d] + ICS_ROOTVSMALL));
            }
            stop = (res->n_iterations >= ctl->max_iter) || (mx < ctl->tolerance) || (mr < ctl->rel_tol);
        }
    } while (!stop);
    // coupledMatrix::solveForIncr: zero the increment in non-solved directions (coupledMatrix.C:371-382)
    for (int d = 0; d < 3; d++)
        if (c->solutionD[d] == -1) {
            LaunchScope ls(c, TM_VEC);
            k_zero_dir<<<g256, 256, 0, c->stream>>>(NP, NPH, d, c->d_x);
        }
    {
        LaunchScope ls(c, TM_VEC);
        k_copy<<<gridFor(V5, 256), 256, 0, c->stream>>>(V5, c->d_x, c->d_dW);
    }
    CUDA_TRY(c, cudaGetLastError());
    return lusgsCheck(c);
}

// ------------------------------------------------------------------------------------------------ C ABI
static int uploadVec5(icsb200_ctx* c, const double* a, const double* b, const double* e, double* dst)
{
    int r;
    if ((r = ics_upload_cells(c, a, 1, dst, c->NPH))) return r;
    if ((r = ics_upload_cells(c, b, 3, dst + c->NPH, c->NPH))) return r;
    return ics_upload_cells(c, e, 1, dst + 4 * (size_t)c->NPH, c->NPH);
}
static int downloadVec5(icsb200_ctx* c, double* a, double* b, double* e, const double* src)
{
    int r = 0;
    if (a && (r = ics_download_cells(c, a, 1, src, c->NPH))) return r;
    if (b && (r = ics_download_cells(c, b, 3, src + c->NPH, c->NPH))) return r;
    if (e && (r = ics_download_cells(c, e, 1, src + 4 * (size_t)c->NPH, c->NPH))) return r;
    return 0;
}

extern "C" int icsb200_matrix_mul(icsb200_ctx* c, const double* xRho, const double* xRhoU, const double* xRhoE, double* yRho, double* yRhoU,
                                  double* yRhoE)
{
    if (!c->matrixSet) return ics_fail(c, ICSB200_ESTATE, "matrix_mul: matrix not assembled");
    int r;
    if ((r = uploadVec5(c, xRho, xRhoU, xRhoE, c->d_x))) return r;
    if ((r = ics_spmv(c, c->d_x, c->d_w, nullptr))) return r;
    return downloadVec5(c, yRho, yRhoU, yRhoE, c->d_w);
}

extern "C" int icsb200_precondition(icsb200_ctx* c, int preconditioner, double* xRho, double* xRhoU, double* xRhoE)
{
    if (!c->matrixSet) return ics_fail(c, ICSB200_ESTATE, "precondition: matrix not assembled");
    int r;
    if ((r = uploadVec5(c, xRho, xRhoU, xRhoE, c->d_x))) return r;
    if ((r = precond(c, preconditioner, c->d_x))) return r;
    if ((r = lusgsCheck(c))) return r;
    return downloadVec5(c, xRho, xRhoU, xRhoE, c->d_x);
}

extern "C" int icsb200_solve_delta(icsb200_ctx* c, const icsb200_solver_controls* ctl, double* dRho, double* dRhoU, double* dRhoE,
                                   icsb200_residuals* res)
{
    if (!c->matrixSet) return ics_fail(c, ICSB200_ESTATE, "solve_delta: matrix not assembled");
    int r = ics_gmres(c, ctl, res);
    if (r) return r;
    c->initRes = *res;
    c->haveInitRes = true;
    if (dRho || dRhoU || dRhoE) r = downloadVec5(c, dRho, dRhoU, dRhoE, c->d_dW);
    else CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return r;
}

// one outer pseudo-time iteration of dbnsFoam: outerLoop.H:51-99 then updateFields.H (dbnsFoam.C:110-124)
static int iterateOnce(icsb200_ctx* c, const icsb200_solver_controls* ctl, icsb200_residuals* res)
{
    int r;
    if ((r = ics_gradients(c))) return r;                 // shared by the flux and the Jacobian reconstructions
    if ((r = ics_flux_residual(c, false))) return r;      // flux.calcFlux + residualsUpdate.H
    if ((r = ics_pseudo_ser(c))) return r;                // setCoAndDeltaT.H (SER)
    if ((r = ics_copy_prev(c))) return r;                 // W -> WPrevIter (outerLoop.H:66-76)
    if ((r = ics_jacobian(c, false))) return r;           // local pseudo dt + createConvectiveJacobian (+ viscous LF)
    if ((r = ics_gmres(c, ctl, res))) return r;           // eqSystem.solveForIncr
    c->initRes = *res;
    c->haveInitRes = true;
    if ((r = ics_bound_local_dt(c))) return r;            // boundLocalTimeStep.H
    if ((r = ics_update(c))) return r;                    // updateFields.H
    c->firstIter = false;
    c->fluxValid = false;
    return 0;
}

extern "C" int icsb200_iterate_dev(icsb200_ctx* c, const icsb200_solver_controls* ctl, icsb200_residuals* res)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "iterate: state not set");
    cudaSetDevice(c->device);
    int r = iterateOnce(c, ctl, res);
    if (r) return r;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int icsb200_iterate_host(icsb200_ctx* c, const icsb200_solver_controls* ctl, double* p, double* U, double* T, icsb200_residuals* res)
{
    if (!c->stateSet) return ics_fail(c, ICSB200_ESTATE, "iterate_host: state not set (call state_set once to initialise boundary data)");
    cudaSetDevice(c->device);
    int r;
    // host fields in: p, U, T of the cells (6N doubles); conserved variables are rebuilt as createFields.H does
    if ((r = ics_upload_cells(c, p, 1, c->q(Q_P), c->NX))) return r;
    if ((r = ics_upload_cells(c, U, 3, c->q(Q_UX), c->NX))) return r;
    if ((r = ics_upload_cells(c, T, 1, c->q(Q_T), c->NX))) return r;
    if ((r = ics_state_from_primitives(c))) return r;
    if ((r = iterateOnce(c, ctl, res))) return r;
    if ((r = ics_download_cells(c, p, 1, c->q(Q_P), c->NX))) return r;
    if ((r = ics_download_cells(c, U, 3, c->q(Q_UX), c->NX))) return r;
    if ((r = ics_download_cells(c, T, 1, c->q(Q_T), c->NX))) return r;
    return 0;
}

// development aid: copy the forward-sweep trace of the last LU-SGS application (ICSB200_LUSGS_TRACE=1) and the slice->level data
extern "C" int icsb200_debug_lusgs_trace(icsb200_ctx* c, long long* out, int* rowNLow, int* cols3)
{
    if (!c->d_lusgsTrace) return ICSB200_ESTATE;
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(out, c->d_lusgsTrace, sizeof(long long) * 4 * c->nSlices, cudaMemcpyDeviceToHost));
    for (int s = 0; s < c->nSlices; s++) {
        rowNLow[s] = c->h_rowNLow[(size_t)s * 32];
        for (int j = 0; j < 3; j++) cols3[3 * s + j] = (j < rowNLow[s]) ? c->h_col[((size_t)c->h_sliceOff[s] + j) * 32] : -1;
    }
    return 0;
}
