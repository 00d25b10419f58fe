// lusgs_blk.cu — LU-SGS sweeps over block tiles (the default schedule; DESIGN.md section 4 "block tiles").
//
// Reference: lusgs::precondition, src/blockFvMatrix/coupledMatrix/preconditioners/lusgs/lusgs.C:220-382 — a forward sweep
// D dW* = R - L dW* followed by a reverse sweep dW = rD (D dW* - U dW), both sequential in cell order.  Any schedule that
// respects the owner < neighbour DAG reproduces the sequential result bit for bit; the level pipeline (solver.cu) walks the
// DAG one hyperplane at a time and pays one L2 round trip per level (3n-2 levels of ~3 us each on an n^3 box), which is
// what bounded the sweeps on small partitions.  Here the mesh is cut into tiles of <= 256 rows (setup.cu: `depth`
// consecutive hyperplanes of an 8x4 column on a structured 3-D mesh, i.e. 8 or 4 levels of 32 rows; 16-wide strips on a
// 2-D mesh) that form a DAG of their own; one CTA sweeps a whole tile out of shared memory and five kinds of warps keep
// every global round trip off the sweep's dependent path:
//   * 10 CONSUMER warps in two groups of five sweep the tile's levels in turn.  A row is swept by FIVE threads (one per
//     component of the 5x5 block row, one warp per component), so the dependent part of a level is 18 shared-memory
//     reads, 30 products and 9 ordered subtractions per thread.  While one group sweeps level L, the other loads everything
//     level L+1 needs that does not depend on the sweep (packed neighbour info, rD, the thread's 15 block coefficients, the
//     right-hand side) into registers; two hardware named barriers, used producer / consumer style, hand the levels over;
//   * the PRODUCER warp streams the 5x5 blocks (600 of ~700 B per row and sweep) with cp.async.bulk (TMA) into a ring of
//     8 slice stages, many slices ahead of the consumers and across tile boundaries;
//   * the METADATA warp deals the tiles: it draws a ticket from a global counter whenever one of the three tile buffers is
//     free and bulk-copies the tile's table (levels, halo positions, flags to wait for), rD, packed row info and (forward)
//     right-hand side into it;
//   * the HALO warp works up to two tiles ahead of the consumers: it polls the epoch flags of the tiles the tile depends on
//     (ld.acquire), gathers the out-of-tile neighbour values into the tile's shared-memory vector and, in the reverse
//     sweep, bulk-copies the tile's own forward values;
//   * the PUBLISH warp waits for the consumers' last sweep of a tile, bulk-stores the tile's vector from shared memory to
//     x, then st.release of the tile's epoch flag, and hands the tile buffer back.  No sentinel buffers: the sweeps run IN
//     PLACE on x exactly like the reference (the forward sweep overwrites the right-hand side with dW* D, the reverse
//     sweep overwrites that with dW).
// Operand order per row is the reference's: neighbours in ascending (forward) / descending (reverse) face order, per
// neighbour the S.S columns (rho, rhoE), then V.S / S.V / V.V (lusgs.C:240-303, 318-380).
#include <algorithm>
#include <climits>
#include <string>

#include "common.cuh"

namespace {

constexpr int MR = ICS_BLK_MR;         // rows per tile
constexpr int MH = ICS_BLK_MH;         // out-of-tile neighbours per tile and sweep
constexpr int XS = MR + MH;            // row stride of the sweep values in shared memory
constexpr int SE = ICS_BLK_SE;         // staged block entries per slice and sweep
constexpr int NCW = 10;                // consumer warps
constexpr int NCT = NCW * 32;          // consumer threads
constexpr int NST = 8;                 // ring stages (a level touches at most ICS_BLK_MAXLW / 32 + 1 = 5 slices)
constexpr int NBUF = 3;                // tile buffers: one being swept, one being prepared by the halo warp, one being written back
constexpr int STAGE_D = SE * 25 * 32;  // doubles per stage
constexpr int NTHREADS = NCT + 128;    // + producer, halo, publish and metadata warps
constexpr int HIT = (MH + 31) / 32;    // halo entries per lane of the halo warp

// a tile's table (setup.cu): 16 descriptor ints, then the sections they point to
enum { BT_T0 = 0, BT_NROWS, BT_COL, BT_NLEV, BT_LEV, BT_HALOF, BT_NHALOF, BT_HALOR, BT_NHALOR, BT_DEPF, BT_NDEPF, BT_DEPR, BT_NDEPR, BT_SLICEOFF, BT_REVLO, BT_NSL };
// profile slots per CTA (ICSB200_LUSGS_PROF), consumer thread 0: cycles waiting for the metadata stage, for the halo warp, in the
// level loops, of which waiting at the level barrier / for block stages; tiles swept, levels swept, total
enum { PF_META = 0, PF_HALO, PF_LEVELS, PF_LVLWAIT, PF_FULLWAIT, PF_TILES, PF_NLEV, PF_TOTAL };

struct BlkArgs {
    int nTiles, nSlices, NP, nCols;
    const int *tab, *idx, *stage, *colStart;
    const unsigned long long* info;
    const double *offd, *rD;
    double* x;
    size_t NPH;
    int* flag;
    int* ticket;  // [2] tile dispenser of this launch (epoch & 1) and of the next one
    int epoch;
    int* err;
    long long* prof;
    long long* trace;  // optional [2 * nTiles][8] global-timer stamps per tile and sweep (ICSB200_LUSGS_TRACE)
};

struct Meta {
    int tab[ICS_BLK_TAB];
    double rD[XS];  // rows of the tile (bulk copy), then the out-of-tile neighbours of the forward sweep (halo warp)
    unsigned long long info[MR];
};

struct BlkSmem {
    double ring[NST][STAGE_D];
    double xs[NBUF][5][XS];  // sweep values: rows of the tile, then the out-of-tile neighbours (forward: already times rD)
    Meta meta[NBUF];
    // the items of this CTA, one per tile buffer: ticket (-1: no more work), the tile's index entry {table offset, table
    // length, first position, rows} and the staging range {first entry, entries} of each of its slices in sweep order
    int itemTicket[NBUF];
    int4 itemIdx[NBUF];
    int2 itemStage[NBUF][MR / 32];
    double zeroB[160];    // the 5x5 block of an absent neighbour (element (k, lane) of a component's row at k*32 + lane)
    unsigned long long full[NST], empty[NST], idfull[NBUF], mfull[NBUF], mempty[NBUF], hfull[NBUF], done[NBUF];
};

__device__ __forceinline__ unsigned sAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbInit(unsigned long long* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sAddr(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbArriveExpectTx(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbExpectTx(unsigned long long* b, unsigned bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(sAddr(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbArrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sAddr(b)) : "memory"); }
// arrival without release semantics: for handing a ring stage back, where the only ordering needed — this warp's reads of the
// stage have completed — holds because their values have been consumed
__device__ __forceinline__ void mbArriveRelaxed(unsigned long long* b)
{
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(sAddr(b)) : "memory");
}
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
// Level hand-over between the two consumer groups: hardware named barriers 1 + g, producer / consumer style — the NCT / 2
// threads of group g arrive (bar.arrive, non-blocking) after sweeping a level, the NCT / 2 threads of the other group wait
// (bar.sync) before sweeping the next one; the barrier completes, and resets, at NCT arrivals.  Strictly alternating by
// construction (a group cannot arrive again before the other group has passed its wait and arrived on the other barrier).
__device__ __forceinline__ void lvlArrive(int g) { asm volatile("bar.arrive %0, %1;" ::"r"(1 + g), "n"(NCT) : "memory"); }
__device__ __forceinline__ void lvlSync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(NCT) : "memory"); }
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sAddr(dstSmem)), "l"(srcGlobal), "r"(bytes),
                 "r"(sAddr(bar))
                 : "memory");
}
__device__ __forceinline__ void bulkStore(void* dstGlobal, const void* srcSmem, unsigned bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(sAddr(srcSmem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ int ldAcquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long gtime()
{
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void stRelease(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// what a consumer thread needs for its (row, component) of a level besides the sweep values of the neighbours
struct Unit {
    int row, lc0, lc1, lc2;
    double rd, xr;
    double B[3][5];
};

template <bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1)
k_lusgs_blk(BlkArgs a)
{
    extern __shared__ __align__(128) unsigned char blkRaw[];
    BlkSmem& sm = *reinterpret_cast<BlkSmem*>(blkRaw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < NST; s++) { mbInit(sm.full + s, 1); mbInit(sm.empty + s, NCW); }
        for (int s = 0; s < NBUF; s++) { mbInit(sm.idfull + s, 1); mbInit(sm.mfull + s, 1); mbInit(sm.mempty + s, 2); mbInit(sm.hfull + s, 1); mbInit(sm.done + s, NCW / 2); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the zero slot absent neighbours point to (last entry of the tile's vector and of its rD), and the zero block
    if (tid < 160) sm.zeroB[tid] = 0.0;
    if (tid < 5 * NBUF) sm.xs[tid / 5][tid % 5][XS - 1] = 0.0;
    if (tid < NBUF) sm.meta[tid].rD[XS - 1] = 0.0;
    __syncthreads();
    // Tiles are dealt dynamically: the metadata warp draws a ticket from a global counter whenever a tile buffer is free —
    // tickets 0 .. nTiles-1 are the tiles of the forward sweep in ascending order, nTiles .. 2 nTiles-1 those of the reverse
    // sweep in descending order — and hands it to the other warps through shared memory.  Tiles are numbered by tile level,
    // every CTA takes its tickets in ascending order and all CTAs are co-resident (cooperative launch), so the holder of the
    // lowest unfinished ticket can always run; a CTA that falls behind simply draws fewer tickets instead of stalling the
    // tiles that a static round-robin deal would have queued behind it.
    const int nT = a.nTiles;
    int* const ticket = a.ticket + (a.epoch & 1);
    if (b == 0 && tid == 0) a.ticket[(a.epoch + 1) & 1] = 0;  // the next launch starts from zero
    // item i of this CTA: wait for its ticket; false = no more work
    auto getItem = [&](int i, bool& fwd, int& tile) {
        const int buf = i % NBUF;
        if (!mbWait(sm.idfull + buf, (i / NBUF) & 1, a.err)) return false;
        const int v = sm.itemTicket[buf];
        if (v < 0) return false;
        fwd = v < nT;
        tile = fwd ? v : 2 * nT - 1 - v;
        return true;
    };

    if (warp == NCW) {
        // ---------------- producer warp: one slice of 5x5 blocks per ring stage, in the order the consumers sweep them, many
        // slices ahead of the consumers and across tile boundaries
        if (lane == 0) {
            int g = 0;
            for (int i = 0;; i++) {
                bool fwd;
                int tile;
                if (!getItem(i, fwd, tile)) break;
                const int buf = i % NBUF;
                const int nSl = sm.itemIdx[buf].w >> 5;
                for (int k = 0; k < nSl; k++, g++) {
                    const int2 sr = sm.itemStage[buf][k];
                    const int st = g % NST;
                    if (g >= NST) {
                        // every consumer has let go of the stage, and its previous fill has landed (rows without neighbours
                        // never wait for a fill, so a stage may be released before its bulk copy completes)
                        const unsigned par = ((g / NST) + 1) & 1;
                        if (!mbWait(sm.empty + st, par, a.err) || !mbWait(sm.full + st, par, a.err)) return;
                    }
                    if (sr.y > 0) {
                        const unsigned bytes = (unsigned)sr.y * 25 * 32 * 8;
                        mbArriveExpectTx(sm.full + st, bytes);
                        bulkLoad(&sm.ring[st][0], a.offd + (size_t)sr.x * 25 * 32, bytes, sm.full + st);
                    } else {
                        mbArrive(sm.full + st);
                    }
                }
            }
        }
        return;
    }

    if (warp == NCW + 3) {
        // ---------------- metadata warp: draws the tickets; a tile's table, rD, packed row info and (forward) right-hand side go
        // into the tile buffer as soon as the publish warp has handed it back, independently of the block ring
        // A ticket is a COLUMN (setup.cu): a run of tiles that this CTA sweeps back to back, each handing its last level to the
        // next one in shared memory; without column mode every tile is a column of its own.
        const int nC = a.nCols;
        int i = 0;
        for (;;) {
            int v = 0;
            if (lane == 0) v = atomicAdd(ticket, 1);
            v = __shfl_sync(0xffffffffu, v, 0);
            if (v >= 2 * nC) {
                const int buf = i % NBUF;
                if (i >= NBUF && !mbWait(sm.mempty + buf, ((i / NBUF) + 1) & 1, a.err)) return;
                if (lane == 0) { sm.itemTicket[buf] = -1; mbArrive(sm.idfull + buf); }
                return;
            }
            const bool fwd = v < nC;
            const int col = fwd ? v : 2 * nC - 1 - v;
            const int c0 = a.colStart[col], c1 = a.colStart[col + 1];
            for (int k = 0; k < c1 - c0; k++, i++) {
                const int tile = fwd ? c0 + k : c1 - 1 - k;
                const int buf = i % NBUF;
                const long long q0 = PROF ? clock64() : 0;
                if (i >= NBUF && !mbWait(sm.mempty + buf, ((i / NBUF) + 1) & 1, a.err)) return;
                if (PROF && lane == 0) a.prof[(size_t)b * 24 + 20] += clock64() - q0;
                const int4 ix = reinterpret_cast<const int4*>(a.idx)[tile];
                const int s0 = ix.z >> 5, nSl = ix.w >> 5;
                if (lane < nSl) {
                    const int sl = fwd ? (s0 + lane) : (s0 + nSl - 1 - lane);
                    sm.itemStage[buf][lane] = reinterpret_cast<const int2*>(a.stage)[(size_t)(fwd ? 0 : a.nSlices) + sl];
                }
                if (lane == 0) { sm.itemTicket[buf] = fwd ? tile : 2 * nT - 1 - tile; sm.itemIdx[buf] = ix; }
                __syncwarp();
                if (lane == 0) {
                    mbArrive(sm.idfull + buf);
                    Meta& M = sm.meta[buf];
                    const unsigned rowB = (unsigned)ix.w * 8, tabB = (unsigned)ix.y * 4;
                    mbArriveExpectTx(sm.mfull + buf, tabB + 2 * rowB + (fwd ? 5 * rowB : 0));
                    bulkLoad(M.tab, a.tab + ix.x, tabB, sm.mfull + buf);
                    bulkLoad(M.info, a.info + (fwd ? (size_t)0 : (size_t)a.NP) + ix.z, rowB, sm.mfull + buf);
                    bulkLoad(M.rD, a.rD + ix.z, rowB, sm.mfull + buf);
                    if (fwd) {
                        for (int kk = 0; kk < 5; kk++) bulkLoad(&sm.xs[buf][kk][0], a.x + kk * a.NPH + ix.z, rowB, sm.mfull + buf);
                    }
                }
                __syncwarp();
            }
        }
    }

    if (warp == NCW + 1) {
        // ---------------- halo warp, one item ahead of the consumers: wait for the tiles this one depends on, then stage the
        // out-of-tile neighbour values (and, reverse sweep, the tile's own forward values) in shared memory
        for (int i = 0;; i++) {
            bool fwd;
            int tile;
            if (!getItem(i, fwd, tile)) return;
            const int buf = i % NBUF;
            Meta& M = sm.meta[buf];
            double(*xs)[XS] = sm.xs[buf];
            long long hk = PROF ? clock64() : 0;
            if (!mbWait(sm.mfull + buf, (i / NBUF) & 1, a.err)) return;
            if (PROF && lane == 0) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 8] += t2 - hk; hk = t2; }
            long long* tr = nullptr;
            if (PROF && a.trace) { tr = a.trace + ((size_t)(fwd ? 0 : nT) + tile) * 8; if (lane == 0) { tr[0] = gtime(); tr[7] = b; } }
            const int* d = M.tab;
            const int t0 = d[BT_T0], nRp = d[BT_NROWS];
            const int* halo = M.tab + (fwd ? d[BT_HALOF] : d[BT_HALOR]);
            const int nHalo = fwd ? d[BT_NHALOF] : d[BT_NHALOR];
            const int* dep = M.tab + (fwd ? d[BT_DEPF] : d[BT_DEPR]);
            const int nDep = fwd ? d[BT_NDEPF] : d[BT_NDEPR];
            int hq[HIT];
#pragma unroll
            for (int it = 0; it < HIT; it++) {
                const int h = it * 32 + lane;
                hq[it] = h < nHalo ? halo[h] : INT_MIN;   // >= 0: global position; -(row) - 1: row of the tile swept just before (column mode)
            }
            // acquire polls, one flag per lane (nothing else of this SM lives in L1, so the invalidate that comes with each costs
            // nothing; a separate fence after relaxed polls, with loads in flight, costs microseconds)
            for (int k = lane; k < nDep; k += 32) {
                const int* f = a.flag + dep[k];
                unsigned int spins = 0;
                while (ldAcquire(f) != a.epoch) {
                    if (++spins > (1u << 24)) { *a.err = 1; break; }
                    if (spins > 64) __nanosleep(32);
                }
            }
            if (PROF && lane == 0) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 9] += t2 - hk; hk = t2; }
            __syncwarp();
            if (PROF && lane == 0) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 10] += t2 - hk; hk = t2; if (tr) tr[1] = gtime(); }
            if (!fwd && lane == 0) {
                // own forward values (written by the CTA that swept this tile forward; covered by the tile's own flag): the
                // acquire above was a generic-proxy read, the bulk copy reads through the async proxy
                asm volatile("fence.proxy.async;" ::: "memory");
                const unsigned rowB = (unsigned)nRp * 8;
                mbExpectTx(sm.hfull + buf, 5 * rowB);
                for (int k = 0; k < 5; k++) bulkLoad(&xs[k][0], a.x + k * a.NPH + t0, rowB, sm.hfull + buf);
            }
            double hv[HIT][5], hsc[HIT];
#pragma unroll
            for (int it = 0; it < HIT; it++) hsc[it] = (fwd && hq[it] >= 0) ? __ldg(a.rD + hq[it]) : 1.0;
#pragma unroll
            for (int it = 0; it < HIT; it++) {
                if (hq[it] >= 0) {
#pragma unroll
                    for (int k = 0; k < 5; k++) hv[it][k] = __ldcg(a.x + k * a.NPH + hq[it]);
                }
            }
#pragma unroll
            for (int it = 0; it < HIT; it++) {
                if (hq[it] >= 0) {
#pragma unroll
                    for (int k = 0; k < 5; k++) xs[k][MR + it * 32 + lane] = hv[it][k];
                    if (fwd) M.rD[MR + it * 32 + lane] = hsc[it];  // the sweep forms dW*_q = rD_q x_q itself (lusgs.C:194-216)
                }
            }
            if ((d[BT_COL] >> (fwd ? 0 : 1)) & 1) {
                // column mode: the previous item of this CTA is the neighbouring chunk of the same column; the rows of it this tile
                // needs come out of its shared-memory vector as soon as its consumers are done — no L2 round trip on the chain
                const int pbuf = (i - 1) % NBUF;
                if (!mbWait(sm.done + pbuf, ((i - 1) / NBUF) & 1, a.err)) return;
                const double(*xp)[XS] = sm.xs[pbuf];
#pragma unroll
                for (int it = 0; it < HIT; it++) {
                    if (hq[it] < 0 && hq[it] != INT_MIN) {
                        const int row = -hq[it] - 1;
#pragma unroll
                        for (int k = 0; k < 5; k++) xs[k][MR + it * 32 + lane] = xp[k][row];
                        if (fwd) M.rD[MR + it * 32 + lane] = sm.meta[pbuf].rD[row];
                    }
                }
            }
            __syncwarp();
            if (PROF && lane == 0) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 11] += t2 - hk; hk = t2; }
            if (PROF && tr && lane == 0) tr[2] = gtime();
            if (lane == 0) {
                mbArrive(sm.hfull + buf);
                if (i > 0) mbArrive(sm.mempty + (i - 1) % NBUF);   // this warp is done with the previous item's buffer, too
            }
        }
        return;
    }

    if (warp == NCW + 2) {
        // ---------------- publish warp: the tile's swept values go from shared memory to x in five bulk stores (no global store
        // sits on the consumers' dependent path; they end their last sweep of the tile with an async-proxy fence); the bulk
        // stores' completion is followed by an implicit generic-async proxy fence, then st.release of the tile's epoch flag (the
        // consumers' arrival on done[] and this thread's release are cumulative) and the tile buffer goes back to the metadata warp.
        if (lane == 0) {
            for (int i = 0;; i++) {
                bool fwd;
                int tile;
                if (!getItem(i, fwd, tile)) return;
                const int buf = i % NBUF;
                long long pk = PROF ? clock64() : 0;
                if (!mbWait(sm.done + buf, (i / NBUF) & 1, a.err)) return;
                if (PROF) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 16] += t2 - pk; pk = t2; }
                const int t0 = sm.meta[buf].tab[BT_T0];
                const unsigned rowB = (unsigned)sm.meta[buf].tab[BT_NROWS] * 8;
                for (int k = 0; k < 5; k++) bulkStore(a.x + k * a.NPH + t0, &sm.xs[buf][k][0], rowB);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                if (PROF) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 17] += t2 - pk; pk = t2; }
                stRelease(a.flag + (fwd ? tile : a.nTiles + tile), a.epoch);
                mbArrive(sm.mempty + buf);
                if (PROF) { const long long t2 = clock64(); a.prof[(size_t)b * 24 + 18] += t2 - pk; pk = t2; if (a.trace) a.trace[((size_t)(fwd ? 0 : a.nTiles) + tile) * 8 + 5] = gtime(); }
            }
        }
        return;
    }

    // ---------------- consumers: two groups of five warps (one warp per component of the block row) take the levels in turn.
    // While one group sweeps level L — the dependent part: neighbour values out of shared memory, 15 products, the ordered
    // subtractions — the other group loads everything level L+1 needs that does not depend on the sweep (packed neighbour
    // info, rD, right-hand side, its 15 block coefficients out of the ring) into registers, so the chain from level to level is
    // barrier -> sweep -> barrier (named barrier 1 + g: group g arrives after each of its sweeps, the other group waits on it).
    const bool prof = PROF && tid == 0;
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long tStart = prof ? clock64() : 0;
    const int r = warp % 5, grp = warp / 5;
    int base = 0;          // ring position of the current item's first slice (all consumers count alike)
    int gl0 = 0;           // levels swept before the current item (both groups count alike): level l of the item belongs to group (gl0 + l) & 1
    for (int i = 0;; i++) {
        bool fwd;
        int tile;
        if (!getItem(i, fwd, tile)) break;
        const int buf = i % NBUF;
        const unsigned bufPar = (i / NBUF) & 1;
        Meta& M = sm.meta[buf];
        const double* xsb = &sm.xs[buf][0][0];
        double* xsr = &sm.xs[buf][r][0];
        long long tk = prof ? clock64() : 0;
        mbWait(sm.mfull + buf, bufPar, a.err);
        if (prof) { const long long t2 = clock64(); pf[PF_META] += t2 - tk; tk = t2; }
        const int* d = M.tab;
        const int nLev = d[BT_NLEV], nSl = d[BT_NSL] & 0xffff;
        const bool unstaged = (d[BT_NSL] >> (fwd ? 16 : 17)) & 1;
        const int* lev = M.tab + d[BT_LEV];
        const int* sliceOff = M.tab + d[BT_SLICEOFF];
        int relNext = 0, relSt = base % NST;  // next slice (in sweep order) and its stage this warp hands back to the producer

        // hand back the stages of the first `target` slices (sweep order) of the tile
        auto releaseTo = [&](int target) {
            __syncwarp();  // every lane's block reads precede lane 0's arrival
            while (relNext < target) {
                if (lane == 0) mbArriveRelaxed(sm.empty + relSt);
                relSt = relSt + 1 == NST ? 0 : relSt + 1;
                relNext++;
            }
        };
        // Everything of unit (chunk rc of level li, component r) that does not depend on the sweep.  Absent neighbours point at
        // the zero slot of the tile's vector and at the zero block (setup.cu), so the sweep itself is branch-free.
        auto loadUnit = [&](Unit& U, int li, int rc) {
            const int L = fwd ? li : nLev - 1 - li;
            const int a0 = lev[L], b0 = lev[L + 1];
            const int row = a0 + rc * 32 + lane;
            U.row = -1;
            if (row < b0) {
                U.row = row;
                const unsigned long long w = M.info[row];
                U.rd = M.rD[row];
                U.xr = xsr[row];
                const int sl = row >> 5, ln = row & 31;
                const int gs = base + (fwd ? sl : nSl - 1 - sl);
                const int st = gs % NST;
                if (PROF) { const long long q0 = prof ? clock64() : 0; mbWait(sm.full + st, (gs / NST) & 1, a.err); if (prof) pf[PF_FULLWAIT] += clock64() - q0; }
                else mbWait(sm.full + st, (gs / NST) & 1, a.err);
                const double* stage = &sm.ring[st][0] + r * 160 + ln;
                const double* zb = sm.zeroB + ln;
#pragma unroll
                for (int t = 0; t < 3; t++) {
                    const int f = (int)(w >> (15 * t)) & 0x7fff;
                    const int lc = f & 1023, code = (f >> 10) & 3;
                    if (t == 0) U.lc0 = lc; else if (t == 1) U.lc1 = lc; else U.lc2 = lc;
                    const double* bp = code == 3 ? zb : stage + code * 800;
                    U.B[t][0] = bp[0]; U.B[t][1] = bp[32]; U.B[t][2] = bp[64]; U.B[t][3] = bp[96]; U.B[t][4] = bp[128];
                }
                if (unstaged) {
                    // rows whose block lies outside the staged entry range of their slice (mesh-boundary slices of the reverse sweep)
#pragma unroll
                    for (int t = 0; t < 3; t++) {
                        const int f = (int)(w >> (15 * t)) & 0x7fff;
                        if (((f >> 10) & 3) == 3 && (f & 1023) != XS - 1) {
                            const double* bp = a.offd + (((size_t)sliceOff[sl] + (f >> 12)) * 25 + r * 5) * 32 + ln;
                            U.B[t][0] = __ldcs(bp); U.B[t][1] = __ldcs(bp + 32); U.B[t][2] = __ldcs(bp + 64); U.B[t][3] = __ldcs(bp + 96); U.B[t][4] = __ldcs(bp + 128);
                        }
                    }
                }
            }
        };
        // the dependent part: neighbour values out of shared memory, ordered subtractions, result to the tile's vector
        auto sweepUnit = [&](const Unit& U) {
            if (U.row < 0) return;
            double xr = U.xr;
#pragma unroll
            for (int t = 0; t < 3; t++) {
                const int lc = t == 0 ? U.lc0 : (t == 1 ? U.lc1 : U.lc2);
                const double* p = xsb + lc;
                double d0 = p[0], d1 = p[XS], d2 = p[2 * XS], d3 = p[3 * XS], d4 = p[4 * XS];
                if (fwd) {  // the lower neighbour's dW* = rD x (lusgs.C:194-216); x itself is what the sweep leaves in place
                    const double rdn = M.rD[lc];
                    d0 = rdn * d0; d1 = rdn * d1; d2 = rdn * d2; d3 = rdn * d3; d4 = rdn * d4;
                }
                // sub-block order of lusgs.C:240-303: S.S (rho column, rhoE column), then the vector columns
                xr -= U.B[t][0] * d0;
                xr -= U.B[t][4] * d4;
                xr -= U.B[t][1] * d1 + U.B[t][2] * d2 + U.B[t][3] * d3;
            }
            xsr[U.row] = fwd ? xr : U.rd * xr;  // forward: the un-scaled running value (lusgs.C:233-237); reverse: dW
        };
        auto chunksOf = [&](int li) { const int L = fwd ? li : nLev - 1 - li; return (lev[L + 1] - lev[L] + 31) >> 5; };

        const int first = ((gl0 & 1) == grp) ? 0 : 1;  // this group's first level of the item
        long long* tr = nullptr;
        if (prof && a.trace) { tr = a.trace + ((size_t)(fwd ? 0 : nT) + tile) * 8; tr[3] = gtime(); }
        if (first < nLev) {
            Unit U;
            // the reverse sweep starts from the tile's own forward values, which the halo warp stages
            if (!fwd) mbWait(sm.hfull + buf, bufPar, a.err);
            loadUnit(U, first, 0);
            if (fwd) mbWait(sm.hfull + buf, bufPar, a.err);
            if (prof) { const long long t2 = clock64(); pf[PF_HALO] += t2 - tk; tk = t2; }
            for (int li = first; li < nLev; li += 2) {
                if (gl0 + li > 0) {   // the level before this one, swept by the other group (possibly as the last level of the previous item)
                    const long long q0 = prof ? clock64() : 0;
                    __syncwarp();
                    lvlSync(grp ^ 1);
                    if (prof) pf[PF_LVLWAIT] += clock64() - q0;
                }
                const long long q2 = prof ? clock64() : 0;
                sweepUnit(U);
                const int nCh = chunksOf(li);
                for (int rc = 1; rc < nCh; rc++) {
                    // levels wider than 32 rows: the remaining chunks, not software-pipelined
                    Unit V;
                    loadUnit(V, li, rc);
                    sweepUnit(V);
                }
                const long long q3 = prof ? clock64() : 0;
                // the bulk stores of the publish warp read what this thread wrote: async-proxy fence after its last sweep of the item
                if (li + 2 >= nLev) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                __threadfence_block();   // bar.arrive itself orders nothing: this thread's stores first
                lvlArrive(grp);
                if (lane == 0 && li == nLev - 1) mbArrive(sm.done + buf);  // every earlier level is ordered before this one through the level barriers
                // The block coefficients of this level have been consumed, so the stages of the slices whose rows all lie in levels
                // <= li go back to the producer now (an arrival right after the loads would wait for them to land instead of
                // letting them overlap the other group's sweep).
                {
                    const int L = fwd ? li : nLev - 1 - li;
                    releaseTo(li == nLev - 1 ? nSl : (fwd ? (lev[L + 1] >> 5) : nSl - ((lev[L] + 31) >> 5)));
                }
                const long long q4 = prof ? clock64() : 0;
                if (li + 2 < nLev) loadUnit(U, li + 2, 0);
                if (prof) { const long long q5 = clock64(); a.prof[(size_t)b * 24 + 12] += q3 - q2; a.prof[(size_t)b * 24 + 15] += q4 - q3; a.prof[(size_t)b * 24 + 13] += q5 - q4; a.prof[(size_t)b * 24 + 21] += 1; }
            }
        }
        releaseTo(nSl);  // slices this group never read from (the other group's last level, or a tile without a level for this group)
        if (prof) { const long long t2 = clock64(); pf[PF_LEVELS] += t2 - tk; tk = t2; pf[PF_TILES]++; pf[PF_NLEV] += nLev; if (tr) tr[4] = gtime(); }
        base += nSl;
        gl0 += nLev;
    }
    if (prof) {
        pf[PF_TOTAL] = clock64() - tStart;
        for (int k = 0; k < 8; k++) a.prof[(size_t)b * 24 + k] = pf[k];
    }
}

}  // namespace

int ics_lusgs_blk(icsb200_ctx* c, double* x)
{
    BlkArgs a{};
    a.nTiles = c->nTiles; a.nSlices = c->nSlices; a.NP = c->NP; a.nCols = c->nBlkCols; a.colStart = c->d_blkCol;
    a.tab = c->d_blkTab; a.idx = c->d_blkIdx; a.stage = c->d_blkStage; a.info = c->d_blkInfo;
    a.offd = c->d_offd; a.rD = c->d_rD; a.x = x; a.NPH = c->NPH;
    a.flag = c->d_blkFlag; a.ticket = c->d_blkFlag + 2 * c->nTiles; a.epoch = ++c->blkEpoch;
    a.err = (int*)c->d_counter + 48;
    const size_t smem = sizeof(BlkSmem) + 128;
    static bool attrSet = false;
    if (!attrSet) {
        CUDA_TRY(c, cudaFuncSetAttribute(k_lusgs_blk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CUDA_TRY(c, cudaFuncSetAttribute(k_lusgs_blk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attrSet = true;
    }
    int grid = std::min(c->numSMs, std::max(1, c->nBlkCols));
    {
        static const char* e3 = getenv("ICSB200_LUSGS_GRID");
        if (e3) grid = std::min(grid, std::max(1, atoi(e3)));
    }
    a.prof = nullptr;
    {
        static const char* e4 = getenv("ICSB200_LUSGS_PROF");
        if (e4) {
            if (!c->d_blkProf) { int r = devAlloc(c, &c->d_blkProf, (size_t)24 * c->numSMs); if (r) return r; }
            a.prof = c->d_blkProf;
            CUDA_TRY(c, cudaMemsetAsync(a.prof, 0, sizeof(long long) * 24 * c->numSMs, c->stream));
        }
    }
    a.trace = nullptr;
    {
        static const char* e5 = getenv("ICSB200_LUSGS_TRACE");
        if (e5 && a.prof) {
            if (!c->d_blkTrace) { int r = devAlloc(c, &c->d_blkTrace, (size_t)16 * c->nTiles); if (r) return r; }
            a.trace = c->d_blkTrace;
        }
    }
    LaunchScope ls(c, TM_LUSGS);
    void* args[] = {&a};
    CUDA_TRY(c, cudaLaunchCooperativeKernel(a.prof ? (void*)k_lusgs_blk<true> : (void*)k_lusgs_blk<false>, dim3(grid), dim3(NTHREADS), args, smem, c->stream));
    return 0;
}

// per-CTA phase cycle counters of the last sweep (development aid, ICSB200_LUSGS_PROF=1): out[grid][24]
// (0..7 consumer thread 0, 8..11 halo warp: metadata wait / flag polls / fence / gather, 16..18 publish warp: wait for the
// consumers / bulk stores / release, 20 metadata warp: wait for a free tile buffer)
extern "C" int icsb200_debug_lusgs_prof(icsb200_ctx* c, long long* out, int max_ctas)
{
    if (!c->d_blkProf) return 0;
    const int n = std::min(max_ctas, std::min(c->numSMs, std::max(1, c->nTiles)));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(out, c->d_blkProf, sizeof(long long) * 24 * n, cudaMemcpyDeviceToHost));
    return n;
}

// global-timer stamps of the last sweep (development aid, ICSB200_LUSGS_PROF=1 ICSB200_LUSGS_TRACE=1): out[2 * nTiles][8] =
// halo warp got the metadata / saw all flags / staged the halo, consumers started / finished, published, -, CTA
extern "C" int icsb200_debug_blk_trace(icsb200_ctx* c, long long* out, int max_rows)
{
    if (!c->d_blkTrace) return 0;
    const int n = std::min(max_rows, 2 * c->nTiles);
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaMemcpy(out, c->d_blkTrace, sizeof(long long) * 8 * n, cudaMemcpyDeviceToHost));
    return n;
}

// the tiles a tile waits for (forward flags), for the trace analysis: out[nTiles][8], -1 padded
extern "C" int icsb200_debug_blk_deps(icsb200_ctx* c, int* out, int max_tiles)
{
    if (!c->blkMode) return 0;
    const int n = std::min(max_tiles, c->nTiles);
    std::vector<int> idx((size_t)4 * c->nTiles);
    CUDA_TRY(c, cudaMemcpy(idx.data(), c->d_blkIdx, sizeof(int) * idx.size(), cudaMemcpyDeviceToHost));
    std::vector<int> tab(ICS_BLK_TAB);
    for (int t = 0; t < n; t++) {
        CUDA_TRY(c, cudaMemcpy(tab.data(), c->d_blkTab + idx[(size_t)4 * t], sizeof(int) * idx[(size_t)4 * t + 1], cudaMemcpyDeviceToHost));
        for (int k = 0; k < 8; k++) out[(size_t)8 * t + k] = k < tab[10] ? tab[tab[9] + k] : -1;
    }
    return n;
}
