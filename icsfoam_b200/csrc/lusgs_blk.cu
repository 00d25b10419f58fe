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
//     with cp.async.bulk (TMA) into a ring of slice stages, many slices ahead of the consumers, across tile boundaries;
//   * the same warp bulk-copies everything else a tile needs that does not depend on the sweep (its table of levels,
//     halo positions and flags to wait for, rD, the packed per-row neighbour info, the right-hand side) into one of two
//     metadata stages while the previous tile is being swept, so a tile starts without a global round trip.
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
constexpr int NST = 7;                 // ring stages (a level touches at most ICS_BLK_MAXLW / 32 + 1 = 5 slices)
constexpr int STAGE_D = SE * 25 * 32;  // doubles per stage

// a tile's table (setup.cu): 16 descriptor ints, then the sections they point to
enum { BT_T0 = 0, BT_NROWS, BT_NREAL, BT_NLEV, BT_LEV, BT_HALOF, BT_NHALOF, BT_HALOR, BT_NHALOR, BT_DEPF, BT_NDEPF, BT_DEPR, BT_NDEPR, BT_SLICEOFF, BT_REVLO, BT_NSL };
// profile slots per CTA (ICSB200_LUSGS_PROF): cycles waiting for the metadata stage, for the flags, loading halo / own values,
// in the level loops, publishing; tiles swept, levels swept, total
enum { PF_META = 0, PF_DEPS, PF_HALO, PF_LEVELS, PF_PUBLISH, PF_TILES, PF_NLEV, PF_TOTAL };

struct BlkArgs {
    int nTiles, nSlices, NP;
    const int *tab, *idx, *stage;
    const unsigned long long* info;
    const double *offd, *rD;
    double* x;
    size_t NPH;
    int* flag;
    int epoch;
    int* err;
    long long* prof;
};

struct Meta {
    int tab[ICS_BLK_TAB];
    double rD[MR];
    unsigned long long info[MR];
};

struct BlkSmem {
    double ring[NST][STAGE_D];
    double xs[2][5][XS];  // sweep values: rows of the tile, then the out-of-tile neighbours (forward: already times rD)
    Meta meta[2];
    unsigned long long full[NST], empty[NST], mfull[2], mempty[2];
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
        for (int s = 0; s < 2; s++) { mbInit(sm.mfull + s, 1); mbInit(sm.mempty + s, 1); }
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
        // ---------------- producer warp.  Per item: the tile's metadata stage (table, rD, row info, forward: right-hand side),
        // then one slice of 5x5 blocks per ring stage in the order the consumers sweep them.  The index entry and the staging
        // ranges of the NEXT item are fetched (lane k: k-th slice) while lane 0 issues the current one.
        int g = 0;
        int4 cur = make_int4(0, 0, 0, 0);
        int curE = 0, curC = 0;
        auto fetch = [&](int i, int4& ix, int& e, int& cnt) {
            bool fwd;
            const int tile = itemTile(i, fwd);
            ix = reinterpret_cast<const int4*>(a.idx)[tile];
            const int s0 = ix.z >> 5, n = ix.w >> 5;
            e = 0; cnt = 0;
            if (lane < n) {
                const int s = fwd ? (s0 + lane) : (s0 + n - 1 - lane);
                const int2 st = reinterpret_cast<const int2*>(a.stage)[(size_t)(fwd ? 0 : a.nSlices) + s];
                e = st.x; cnt = st.y;
            }
        };
        if (nItems > 0) fetch(0, cur, curE, curC);
        for (int i = 0; i < nItems; i++) {
            int4 nx = make_int4(0, 0, 0, 0);
            int nxE = 0, nxC = 0;
            if (i + 1 < nItems) fetch(i + 1, nx, nxE, nxC);
            const bool fwd = i < nMine;
            const int nSl = cur.w >> 5;
            if (lane == 0) {
                const int buf = i & 1;
                bool okw = true;
                if (i >= 2) okw = mbWait(sm.mempty + buf, ((i >> 1) + 1) & 1, a.err);
                if (okw) {
                    Meta& M = sm.meta[buf];
                    const unsigned rowB = (unsigned)cur.w * 8, tabB = (unsigned)cur.y * 4;
                    mbExpectTx(sm.mfull + buf, tabB + 2 * rowB + (fwd ? 5 * rowB : 0));
                    bulkLoad(M.tab, a.tab + cur.x, tabB, sm.mfull + buf);
                    bulkLoad(M.info, a.info + (fwd ? (size_t)0 : (size_t)a.NP) + cur.z, rowB, sm.mfull + buf);
                    bulkLoad(M.rD, a.rD + cur.z, rowB, sm.mfull + buf);
                    if (fwd) {
                        for (int k = 0; k < 5; k++) bulkLoad(&sm.xs[buf][k][0], a.x + k * a.NPH + cur.z, rowB, sm.mfull + buf);
                    }
                }
            }
            for (int k = 0; k < nSl; k++) {
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
            cur = nx; curE = nxE; curC = nxC;
        }
        return;
    }

    // ---------------- consumers ----------------
    const bool prof = a.prof != nullptr && tid == 0;
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tStart = prof ? clock64() : 0;
    int base = 0;  // ring position of the current item's first slice (all consumers count alike)
    for (int i = 0; i < nItems; i++) {
        bool fwd;
        const int tile = itemTile(i, fwd);
        const int buf = i & 1;
        Meta& M = sm.meta[buf];
        double(*xs)[XS] = sm.xs[buf];
        long long tk = prof ? clock64() : 0;
        mbWait(sm.mfull + buf, (i >> 1) & 1, a.err);
        if (prof) { const long long t2 = clock64(); pf[PF_META] += t2 - tk; tk = t2; }
        const int* d = M.tab;
        const int t0 = d[BT_T0], nRp = d[BT_NROWS], nLev = d[BT_NLEV], nSl = d[BT_NSL];
        const int* lev = M.tab + d[BT_LEV];
        const int* halo = M.tab + (fwd ? d[BT_HALOF] : d[BT_HALOR]);
        const int nHalo = fwd ? d[BT_NHALOF] : d[BT_NHALOR];
        const int* dep = M.tab + (fwd ? d[BT_DEPF] : d[BT_DEPR]);
        const int nDep = fwd ? d[BT_NDEPF] : d[BT_NDEPR];
        const int* sliceOff = M.tab + d[BT_SLICEOFF];
        // out-of-tile neighbours of this sweep (nHalo <= MH <= NCT): position and scale before the wait, value after it
        int hq = -1;
        double hsc = 1.0;
        if (tid < nHalo) { hq = halo[tid]; if (fwd) hsc = a.rD[hq]; }
        // wait for the tiles this one depends on (one flag per thread)
        if (tid < nDep) {
            const int* f = a.flag + dep[tid];
            unsigned int spins = 0;
            while (ldAcquire(f) != a.epoch) {
                if (++spins > (1u << 24)) { *a.err = 1; break; }
                if (spins > 32) __nanosleep(64);
            }
        }
        consumerBarrier();
        if (prof) { const long long t2 = clock64(); pf[PF_DEPS] += t2 - tk; tk = t2; }
        {
            // reverse: own forward values (written by the CTA that swept this tile forward; covered by the tile's own flag).
            // All loads of the thread are issued before the first store.
            double own[2][5], hv[5];
            const int r1 = tid, r2 = tid + NCT;
            if (!fwd) {
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    own[0][k] = (r1 < nRp) ? __ldcg(a.x + k * a.NPH + t0 + r1) : 0.0;
                    own[1][k] = (r2 < nRp) ? __ldcg(a.x + k * a.NPH + t0 + r2) : 0.0;
                }
            }
            if (hq >= 0) {
#pragma unroll
                for (int k = 0; k < 5; k++) hv[k] = __ldcg(a.x + k * a.NPH + hq);
            }
            if (!fwd) {
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    if (r1 < nRp) xs[k][r1] = own[0][k];
                    if (r2 < nRp) xs[k][r2] = own[1][k];
                }
            }
            if (hq >= 0) {
#pragma unroll
                for (int k = 0; k < 5; k++) xs[k][MR + tid] = fwd ? hsc * hv[k] : hv[k];  // dW*_q = rD_q x_q (lusgs.C:194-216)
            }
        }
        consumerBarrier();
        if (prof) { const long long t2 = clock64(); pf[PF_HALO] += t2 - tk; tk = t2; }

        // ---- the tile's levels: forward ascending, reverse descending (rows of one forward level never depend on each other,
        // and every upper neighbour sits in a higher forward level, so the forward levels are a valid reverse schedule too)
        int relNext = 0;  // thread 0: next slice (in sweep order) whose stage can be handed back to the producer
        for (int li = 0; li < nLev; li++) {
            const int L = fwd ? li : nLev - 1 - li;
            const int a0 = lev[L], b0 = lev[L + 1];
            const int nUnits = 5 * ((b0 - a0 + 31) >> 5);
            for (int u = warp; u < nUnits; u += NCW) {
                const int rc = u / 5, r = u - rc * 5;  // unit = (chunk of 32 rows, component r of the block row)
                const int row = a0 + rc * 32 + lane;
                if (row < b0) {
                    const unsigned long long w = M.info[row];
                    const int n = (int)(w >> 45) & 3;
                    double xr = xs[r][row];
                    const int sl = row >> 5, ln = row & 31;
                    const int gs = base + (fwd ? sl : nSl - 1 - sl);
                    const int st = gs % NST;
                    if (n > 0) mbWait(sm.full + st, (gs / NST) & 1, a.err);
                    const double* stage = &sm.ring[st][0] + r * 5 * 32 + ln;
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        if (t < n) {
                            const int f = (int)(w >> (15 * t)) & 0x7fff;
                            const int lc = f & 1023, code = (f >> 10) & 3, j = f >> 12;
                            const double d0 = xs[0][lc], d1 = xs[1][lc], d2 = xs[2][lc], d3 = xs[3][lc], d4 = xs[4][lc];
                            double B0, B1, B2, B3, B4;
                            if (code < 3) {
                                const double* bp = stage + code * 25 * 32;
                                B0 = bp[0]; B1 = bp[32]; B2 = bp[64]; B3 = bp[96]; B4 = bp[128];
                            } else {
                                const double* bp = a.offd + (((size_t)sliceOff[sl] + j) * 25 + r * 5) * 32 + ln;
                                B0 = __ldcs(bp); B1 = __ldcs(bp + 32); B2 = __ldcs(bp + 64); B3 = __ldcs(bp + 96); B4 = __ldcs(bp + 128);
                            }
                            // sub-block order of lusgs.C:240-303: S.S (rho column, rhoE column), then the vector columns
                            xr -= B0 * d0;
                            xr -= B4 * d4;
                            xr -= B1 * d1 + B2 * d2 + B3 * d3;
                        }
                    }
                    const double rd = M.rD[row];
                    if (fwd) {
                        __stcg(a.x + r * a.NPH + t0 + row, xr);  // un-scaled running value (lusgs.C:233-237)
                        xs[r][row] = rd * xr;                    // what the upper neighbours subtract: rD x
                    } else {
                        const double v = rd * xr;
                        __stcg(a.x + r * a.NPH + t0 + row, v);
                        xs[r][row] = v;
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
        if (prof) { const long long t2 = clock64(); pf[PF_LEVELS] += t2 - tk; tk = t2; pf[PF_TILES]++; pf[PF_NLEV] += nLev; }
        // publish: every consumer's stores precede the last level barrier; the fence + release by one thread is cumulative.
        // Every consumer is past its last read of the metadata stage, so it goes back to the producer as well.
        if (tid == 0) {
            __threadfence();
            stRelease(a.flag + (fwd ? tile : a.nTiles + tile), a.epoch);
            mbArrive(sm.mempty + buf);
        }
        if (prof) { const long long t2 = clock64(); pf[PF_PUBLISH] += t2 - tk; }
        base += nSl;
    }
    if (prof) {
        pf[PF_TOTAL] = clock64() - tStart;
        for (int k = 0; k < 8; k++) a.prof[(size_t)b * 8 + k] = pf[k];
    }
}

}  // namespace

int ics_lusgs_blk(icsb200_ctx* c, double* x)
{
    BlkArgs a{};
    a.nTiles = c->nTiles; a.nSlices = c->nSlices; a.NP = c->NP;
    a.tab = c->d_blkTab; a.idx = c->d_blkIdx; a.stage = c->d_blkStage; a.info = c->d_blkInfo;
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
    a.prof = nullptr;
    {
        static const char* e4 = getenv("ICSB200_LUSGS_PROF");
        if (e4) {
            if (!c->d_blkProf) { int r = devAlloc(c, &c->d_blkProf, (size_t)8 * c->numSMs); if (r) return r; }
            a.prof = c->d_blkProf;
        }
    }
    LaunchScope ls(c, TM_LUSGS);
    void* args[] = {&a};
    CUDA_TRY(c, cudaLaunchCooperativeKernel((void*)k_lusgs_blk, dim3(grid), dim3(NCT + 32), args, smem, c->stream));
    return 0;
}

// per-CTA phase cycle counters of the last sweep (development aid, ICSB200_LUSGS_PROF=1): out[grid][8]
extern "C" int icsb200_debug_lusgs_prof(icsb200_ctx* c, long long* out, int max_ctas)
{
    if (!c->d_blkProf) return 0;
    const int n = std::min(max_ctas, std::min(c->numSMs, std::max(1, c->nTiles)));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(out, c->d_blkProf, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost));
    return n;
}
