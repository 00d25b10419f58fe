// lusgs_blk.cu — LU-SGS sweeps over block tiles (the default schedule; DESIGN.md section 4 "block tiles").
//
// Reference: lusgs::precondition, src/blockFvMatrix/coupledMatrix/preconditioners/lusgs/lusgs.C:220-382 — a forward sweep
// D dW* = R - L dW* followed by a reverse sweep dW = rD (D dW* - U dW), both sequential in cell order.  Any schedule that
// respects the owner < neighbour DAG reproduces the sequential result bit for bit; the level pipeline (solver.cu) walks the
// DAG one hyperplane at a time and pays one L2 round trip per level (3n-2 levels of ~3 us each on an n^3 box), which is
// what bounded the sweeps on small partitions.  Here the mesh is cut into tiles of <= 512 rows (8x8x8 cells on a
// structured mesh, setup.cu) that form a DAG of their own; one CTA sweeps a whole tile out of shared memory:
//   * dependency hops INSIDE a tile cost a named barrier + shared-memory reads (~0.1 us) instead of an L2 round trip;
//   * hops BETWEEN tiles (3n/8 - 2 of them) use an epoch flag per tile and sweep, published with fence + st.release and
//     polled with ld.acquire — no sentinel buffers, so the sweeps run IN PLACE on x exactly like the reference: the
//     forward sweep overwrites the right-hand side with dW* D, the reverse sweep overwrites that with dW;
//   * a row is swept by FIVE threads (one per component of the 5x5 block row), so the dependent part of a level is
//     15 products + 9 ordered subtractions per thread instead of 75 + 45;
//   * the 5x5 blocks (all of the traffic that matters: 600 of ~700 B per row and sweep) are streamed by a producer warp
//     with cp.async.bulk (TMA) into a ring of slice stages, many slices ahead of the consumers, across tile boundaries.
// Operand order per row is the reference's: neighbours in ascending (forward) / descending (reverse) face order, per
// neighbour the S.S columns (rho, rhoE), then V.S / S.V / V.V (lusgs.C:240-303, 318-380).
#include <algorithm>
#include <string>

#include "common.cuh"

namespace {

constexpr int MR = ICS_BLK_MR;         // rows per tile
constexpr int MH = ICS_BLK_MH;         // out-of-tile neighbours per tile and sweep
constexpr int XS = MR + MH;            // row stride of the sweep values in shared memory
constexpr int SE = ICS_BLK_SE;         // staged block entries per slice and sweep
constexpr int NCW = 10;                // consumer warps
constexpr int NCT = NCW * 32;          // consumer threads
constexpr int NST = 9;                 // ring stages
constexpr int STAGE_D = SE * 25 * 32;  // doubles per stage
constexpr int MAXSL = MR / 32;         // slices per tile

enum { BD_T0 = 0, BD_NROWS, BD_NREAL, BD_NLEV, BD_LEVPTR, BD_HALOF, BD_NHALOF, BD_HALOR, BD_NHALOR, BD_DEPF, BD_NDEPF, BD_DEPR, BD_NDEPR };

struct BlkArgs {
    int nTiles, nSlices, NP;
    const int *desc, *levTab, *halo, *dep, *stage, *sliceOff, *sliceRevLo, *rowNLow, *rowNInt;
    const short* lcol;
    const double *offd, *rD;
    double* x;
    size_t NPH;
    int* flag;
    int epoch;
    int* err;
};

struct BlkSmem {
    double ring[NST][STAGE_D];
    double xs[5][XS];       // sweep values: rows of the tile, then the out-of-tile neighbours (forward: already times rD)
    double rD[MR];
    int nli[MR];            // nLow | nInt << 8
    short lcol[3][MR];      // this sweep's local neighbour indices
    int lev[ICS_BLK_MAXLEV + 1];
    int stageLo[MAXSL];     // first staged entry of a slice (relative to the slice's first entry)
    int sliceOff[MAXSL];    // first entry of a slice (global fallback for entries that are not staged)
    unsigned long long full[NST], empty[NST];
};

__device__ __forceinline__ unsigned sAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbInit(unsigned long long* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sAddr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbExpectTx(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbArrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sAddr(b)) : "memory"); }
__device__ __forceinline__ bool mbTry(unsigned long long* b, unsigned parity)
{
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sAddr(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbWait(unsigned long long* b, unsigned parity, int* err)
{
    unsigned int spins = 0;
    while (!mbTry(b, parity)) {
        if (++spins > (1u << 26)) { *err = 2; return false; }
    }
    return true;
}
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sAddr(dstSmem)), "l"(srcGlobal), "r"(bytes),
                 "r"(sAddr(bar))
                 : "memory");
}
__device__ __forceinline__ int ldAcquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelease(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void consumerBarrier() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

__global__ void __launch_bounds__(NCT + 32, 1)
k_lusgs_blk(BlkArgs a)
{
    extern __shared__ __align__(128) unsigned char blkRaw[];
    BlkSmem& sm = *reinterpret_cast<BlkSmem*>(blkRaw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = gridDim.x, b = blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbInit(sm.full + s, 1); mbInit(sm.empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // items of this CTA: tiles b, b+G, ... ascending (forward sweep), then nTiles-1-b, nTiles-1-b-G, ... (reverse sweep).
    // Tiles are numbered by tile level, so neighbouring tiles of a level go to different SMs, every CTA walks its items in
    // dependency order and — all CTAs being co-resident (cooperative launch) — the lowest unfinished tile can always run.
    const int nMine = (a.nTiles > b) ? (a.nTiles - b + G - 1) / G : 0;
    const int nItems = 2 * nMine;
    auto itemTile = [&](int i, bool& fwd) { fwd = i < nMine; return fwd ? (b + i * G) : (a.nTiles - 1 - (b + (i - nMine) * G)); };

    if (warp == NCW) {
        // ---------------- producer warp: one slice of 5x5 blocks per ring stage, in the order the consumers sweep them.
        // Lane k fetches the staging range of the k-th slice of the NEXT item while lane 0 issues the current one.
        int g = 0;
        int curS0 = 0, curN = 0, curE = 0, curC = 0;
        auto fetch = [&](int i, int& s0, int& n, int& e, int& cnt) {
            bool fwd;
            const int tile = itemTile(i, fwd);
            const int t0 = a.desc[(size_t)16 * tile + BD_T0], nr = a.desc[(size_t)16 * tile + BD_NROWS];
            s0 = t0 >> 5; n = nr >> 5;
            e = 0; cnt = 0;
            if (lane < n) {
                const int s = fwd ? (s0 + lane) : (s0 + n - 1 - lane);
                const int2 st = reinterpret_cast<const int2*>(a.stage)[(size_t)(fwd ? 0 : a.nSlices) + s];
                e = st.x; cnt = st.y;
            }
        };
        if (nItems > 0) fetch(0, curS0, curN, curE, curC);
        for (int i = 0; i < nItems; i++) {
            int nxS0 = 0, nxN = 0, nxE = 0, nxC = 0;
            if (i + 1 < nItems) fetch(i + 1, nxS0, nxN, nxE, nxC);
            for (int k = 0; k < curN; k++) {
                const int e = __shfl_sync(0xffffffffu, curE, k), cnt = __shfl_sync(0xffffffffu, curC, k);
                if (lane == 0) {
                    const int st = g % NST;
                    bool okw = true;
                    if (g >= NST) okw = mbWait(sm.empty + st, ((g / NST) + 1) & 1, a.err);
                    if (okw) {
                        if (cnt > 0) {
                            const unsigned bytes = (unsigned)cnt * 25 * 32 * 8;
                            mbExpectTx(sm.full + st, bytes);
                            bulkLoad(&sm.ring[st][0], a.offd + (size_t)e * 25 * 32, bytes, sm.full + st);
                        } else {
                            mbArrive(sm.full + st);
                        }
                    }
                }
                g++;
            }
            curS0 = nxS0; curN = nxN; curE = nxE; curC = nxC;
        }
        return;
    }

    // ---------------- consumers ----------------
    int base = 0;  // ring position of the current item's first slice (all consumers count alike)
    for (int i = 0; i < nItems; i++) {
        bool fwd;
        const int tile = itemTile(i, fwd);
        const int* d = a.desc + (size_t)16 * tile;
        const int t0 = d[BD_T0], nRp = d[BD_NROWS], nLev = d[BD_NLEV], levPtr = d[BD_LEVPTR];
        const int haloPtr = fwd ? d[BD_HALOF] : d[BD_HALOR], nHalo = fwd ? d[BD_NHALOF] : d[BD_NHALOR];
        const int depPtr = fwd ? d[BD_DEPF] : d[BD_DEPR], nDep = fwd ? d[BD_NDEPF] : d[BD_NDEPR];
        const int nSl = nRp >> 5, s0 = t0 >> 5;
        // everything that does not depend on the sweep: row metadata, local neighbour indices, level table, staging offsets,
        // and (forward) the right-hand side
        for (int r = tid; r < nRp; r += NCT) {
            const int p = t0 + r;
            sm.rD[r] = a.rD[p];
            sm.nli[r] = a.rowNLow[p] | (a.rowNInt[p] << 8);
#pragma unroll
            for (int t = 0; t < 3; t++) sm.lcol[t][r] = a.lcol[(size_t)((fwd ? 0 : 3) + t) * a.NP + p];
            if (fwd) {
#pragma unroll
                for (int k = 0; k < 5; k++) sm.xs[k][r] = __ldcg(a.x + k * a.NPH + p);
            }
        }
        for (int L = tid; L <= nLev; L += NCT) sm.lev[L] = a.levTab[levPtr + L];
        for (int s = tid; s < nSl; s += NCT) { sm.stageLo[s] = fwd ? 0 : a.sliceRevLo[s0 + s]; sm.sliceOff[s] = a.sliceOff[s0 + s]; }
        // out-of-tile neighbours of this sweep (nHalo <= MH <= NCT): index and scale before the wait, value after it
        int hq = -1;
        double hsc = 1.0;
        if (tid < nHalo) { hq = a.halo[haloPtr + tid]; if (fwd) hsc = a.rD[hq]; }
        // wait for the tiles this one depends on (one flag per thread)
        if (tid < nDep) {
            const int* f = a.flag + a.dep[depPtr + tid];
            unsigned int spins = 0;
            while (ldAcquire(f) != a.epoch) {
                if (++spins > (1u << 24)) { *a.err = 1; break; }
                if (spins > 32) __nanosleep(64);
            }
        }
        consumerBarrier();
        if (!fwd) {  // own forward values (written by the CTA that swept this tile forward; covered by the tile's own flag)
            for (int r = tid; r < nRp; r += NCT) {
#pragma unroll
                for (int k = 0; k < 5; k++) sm.xs[k][r] = __ldcg(a.x + k * a.NPH + t0 + r);
            }
        }
        if (hq >= 0) {
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const double v = __ldcg(a.x + k * a.NPH + hq);
                sm.xs[k][MR + tid] = fwd ? hsc * v : v;  // dW*_q = rD_q x_q (lusgs.C:194-216)
            }
        }
        consumerBarrier();

        // ---- the tile's levels: forward ascending, reverse descending (rows of one forward level never depend on each other,
        // and every upper neighbour sits in a higher forward level, so the forward levels are a valid reverse schedule too)
        int relNext = 0;  // thread 0: next slice (in sweep order) whose stage can be handed back to the producer
        for (int li = 0; li < nLev; li++) {
            const int L = fwd ? li : nLev - 1 - li;
            const int a0 = sm.lev[L], b0 = sm.lev[L + 1];
            const int nUnits = 5 * ((b0 - a0 + 31) >> 5);
            for (int u = warp; u < nUnits; u += NCW) {
                const int rc = u / 5, r = u - rc * 5;  // unit = (chunk of 32 rows, component r of the block row)
                const int row = a0 + rc * 32 + lane;
                if (row < b0) {
                    const int m = sm.nli[row];
                    const int nLow = m & 255, nInt = m >> 8;
                    const int n = fwd ? nLow : nInt - nLow;
                    double xr = sm.xs[r][row];
                    const int sl = row >> 5, ln = row & 31;
                    const int gs = base + (fwd ? sl : nSl - 1 - sl);
                    const int st = gs % NST;
                    const int lo = sm.stageLo[sl];
                    if (n > 0) mbWait(sm.full + st, (gs / NST) & 1, a.err);
                    const double* stage = &sm.ring[st][0] + ln;
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        if (t < n) {
                            const int j = fwd ? t : nInt - 1 - t;
                            const int lc = sm.lcol[t][row];
                            const double d0 = sm.xs[0][lc], d1 = sm.xs[1][lc], d2 = sm.xs[2][lc], d3 = sm.xs[3][lc], d4 = sm.xs[4][lc];
                            const int js = j - lo;
                            double B0, B1, B2, B3, B4;
                            if (js >= 0 && js < SE) {
                                const double* bp = stage + (size_t)(js * 25 + r * 5) * 32;
                                B0 = bp[0]; B1 = bp[32]; B2 = bp[64]; B3 = bp[96]; B4 = bp[128];
                            } else {
                                const double* bp = a.offd + (((size_t)sm.sliceOff[sl] + j) * 25 + r * 5) * 32 + ln;
                                B0 = __ldcs(bp); B1 = __ldcs(bp + 32); B2 = __ldcs(bp + 64); B3 = __ldcs(bp + 96); B4 = __ldcs(bp + 128);
                            }
                            // sub-block order of lusgs.C:240-303: S.S (rho column, rhoE column), then the vector columns
                            xr -= B0 * d0;
                            xr -= B4 * d4;
                            xr -= B1 * d1 + B2 * d2 + B3 * d3;
                        }
                    }
                    const double rd = sm.rD[row];
                    if (fwd) {
                        __stcg(a.x + r * a.NPH + t0 + row, xr);  // un-scaled running value (lusgs.C:233-237)
                        sm.xs[r][row] = rd * xr;                 // what the upper neighbours subtract: rD x
                    } else {
                        const double v = rd * xr;
                        __stcg(a.x + r * a.NPH + t0 + row, v);
                        sm.xs[r][row] = v;
                    }
                }
            }
            consumerBarrier();
            if (tid == 0) {
                const bool last = li == nLev - 1;
                while (relNext < nSl) {
                    const int sl = fwd ? relNext : nSl - 1 - relNext;
                    const bool done = last || (fwd ? ((sl + 1) * 32 <= b0) : (sl * 32 >= a0));
                    if (!done) break;
                    // the fill of this round must have landed before the stage is handed back (rows without neighbours never waited)
                    mbWait(sm.full + (base + relNext) % NST, ((base + relNext) / NST) & 1, a.err);
                    mbArrive(sm.empty + (base + relNext) % NST);
                    relNext++;
                }
            }
        }
        // publish: every consumer's stores precede the barrier above; the fence + release by one thread is cumulative
        if (tid == 0) {
            __threadfence();
            stRelease(a.flag + (fwd ? tile : a.nTiles + tile), a.epoch);
        }
        base += nSl;
        // the next item overwrites the staged metadata: nobody may still be reading it (all are past the last level barrier)
    }
}

}  // namespace

int ics_lusgs_blk(icsb200_ctx* c, double* x)
{
    BlkArgs a{};
    a.nTiles = c->nTiles; a.nSlices = c->nSlices; a.NP = c->NP;
    a.desc = c->d_blkDesc; a.levTab = c->d_tileFLev; a.halo = c->d_blkHalo; a.dep = c->d_blkDep; a.stage = c->d_blkStage;
    a.sliceOff = c->d_sliceOff; a.sliceRevLo = c->d_sliceRange + c->nSlices; a.rowNLow = c->d_rowNLow; a.rowNInt = c->d_rowNInt;
    a.lcol = c->d_blkLcol;
    a.offd = c->d_offd; a.rD = c->d_rD; a.x = x; a.NPH = c->NPH;
    a.flag = c->d_blkFlag; a.epoch = ++c->blkEpoch;
    a.err = (int*)c->d_counter + 48;
    const size_t smem = sizeof(BlkSmem) + 128;
    static bool attrSet = false;
    if (!attrSet) {
        CUDA_TRY(c, cudaFuncSetAttribute(k_lusgs_blk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attrSet = true;
    }
    int grid = std::min(c->numSMs, std::max(1, c->nTiles));
    {
        static const char* e3 = getenv("ICSB200_LUSGS_GRID");
        if (e3) grid = std::min(grid, std::max(1, atoi(e3)));
    }
    LaunchScope ls(c, TM_LUSGS);
    void* args[] = {&a};
    CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)k_lusgs_blk, dim3(grid), dim3(NCT + 32), args, smem, c->stream));
    return 0;
}
