// setup.cu — context lifetime, mesh ingestion (level schedule, positions, sliced-ELL rows), selectors.
//
// Replaces, on the reference side: the lduAddressing / mesh.cells() traversal order the solver relies on
// (lusgs.C:141-156, blockFvMatrix.C:340-380), construction of coupledMatrix storage (coupledMatrix.C:45-61,
// re-done every outer iteration at outerLoop.H:53 — here allocated once), and run-time selection
// (newConvectiveFluxScheme.C:39-69, coupledMatrixSolver.C:41-67, coupledMatrixPreconditioner.C:41-64).
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------ kernels
__global__ void k_gather_cells(const double* __restrict__ stage, int nc, const int* __restrict__ pos2cell, int NP, double* __restrict__ dst,
                               size_t stride, double padValue)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    int c = pos2cell[p];
    for (int k = 0; k < nc; k++) dst[k * stride + p] = (c >= 0) ? stage[(size_t)c * nc + k] : padValue;
}

__global__ void k_scatter_cells(double* __restrict__ stage, int nc, const int* __restrict__ pos2cell, int NP, const double* __restrict__ src,
                                size_t stride)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= NP) return;
    int c = pos2cell[p];
    if (c < 0) return;
    for (int k = 0; k < nc; k++) stage[(size_t)c * nc + k] = src[k * stride + p];
}

int ics_ensure_stage(icsb200_ctx* c, size_t bytes)
{
    if (c->stageBytes >= bytes) return 0;
    if (c->d_stage) cudaFree(c->d_stage);
    c->d_stage = nullptr;
    c->stageBytes = 0;
    CUDA_TRY(c, cudaMalloc((void**)&c->d_stage, bytes));
    c->stageBytes = bytes;
    return 0;
}

int ics_upload_cells(icsb200_ctx* c, const double* host, int nc, double* dst, size_t stride)
{
    size_t bytes = (size_t)c->N * nc * sizeof(double);
    int r = ics_ensure_stage(c, bytes);
    if (r) return r;
    CUDA_TRY(c, cudaMemcpyAsync(c->d_stage, host, bytes, cudaMemcpyHostToDevice, c->stream));
    {
        LaunchScope ls(c, TM_PERM);
        k_gather_cells<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->d_stage, nc, c->d_pos2cell, c->NP, dst, stride, 0.0);
    }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
}

int ics_download_cells(icsb200_ctx* c, double* host, int nc, const double* src, size_t stride)
{
    size_t bytes = (size_t)c->N * nc * sizeof(double);
    int r = ics_ensure_stage(c, bytes);
    if (r) return r;
    {
        LaunchScope ls(c, TM_PERM);
        k_scatter_cells<<<gridFor(c->NP, 256), 256, 0, c->stream>>>(c->d_stage, nc, c->d_pos2cell, c->NP, src, stride);
    }
    CUDA_TRY(c, cudaGetLastError());
    CUDA_TRY(c, cudaMemcpyAsync(host, c->d_stage, bytes, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------ lifetime
extern "C" int icsb200_nccl_unique_id(void* out128)
{
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return ICSB200_ECUDA;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    std::memcpy(out128, &id, 128);
    return 0;
}

extern "C" int icsb200_create(icsb200_ctx** out, int device, const void* nccl_unique_id, int rank, int n_ranks)
{
    *out = nullptr;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0) return ICSB200_ECUDA;  // no CPU fallback
    if (device < 0 || device >= nDev) return ICSB200_EINVAL;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ICSB200_ECUDA;
    if (prop.major != 10) return ICSB200_ECUDA;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return ICSB200_ECUDA;
    icsb200_ctx* c = new icsb200_ctx;
    c->device = device;
    c->rank = rank;
    c->nRanks = n_ranks;
    c->numSMs = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->commStream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return ICSB200_ECUDA; }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaEventCreate(&c->evA);
    cudaEventCreate(&c->evB);
    cudaMalloc((void**)&c->d_scal, 4096 * sizeof(double));
    cudaMallocHost((void**)&c->h_scal, 4096 * sizeof(double));
    cudaMalloc((void**)&c->d_partial, 64 * 4096 * sizeof(double));
    cudaMalloc((void**)&c->d_counter, 64 * sizeof(unsigned int));
    cudaMemset(c->d_counter, 0, 64 * sizeof(unsigned int));
    cudaMalloc((void**)&c->d_barrier, 64 * sizeof(unsigned int));
    cudaMemset(c->d_barrier, 0, 64 * sizeof(unsigned int));
    if (n_ranks > 1) {
        if (!nccl_unique_id) { delete c; return ICSB200_EINVAL; }
        ncclUniqueId id;
        std::memcpy(&id, nccl_unique_id, 128);
        ncclComm_t comm;
        if (ncclCommInitRank(&comm, n_ranks, id, rank) != ncclSuccess) { delete c; return ICSB200_ECUDA; }
        c->nccl = comm;
    }
    *out = c;
    return 0;
}

extern "C" int icsb200_destroy(icsb200_ctx* c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->nccl) ncclCommDestroy((ncclComm_t)c->nccl);
    void* ptrs[] = {c->d_pos2cell, c->d_cell2pos, c->d_sliceOff, c->d_rowNLow, c->d_rowNInt, c->d_rowNAll, c->d_col, c->d_meta, c->d_gfid,
                    c->d_geo, c->d_dCoupled, c->d_V, c->d_C, c->d_tileStart, c->d_tileFPtr, c->d_tileFLev, c->d_tileRPtr, c->d_tileRLev, c->d_tileRRows, c->d_sliceTile, c->d_bfOwnerPos, c->d_bfPatch, c->d_bfKind,
                    c->d_bfGeo, c->d_bc, c->d_phiB, c->d_vic, c->d_sendBuf, c->d_recvBuf, c->d_fields, c->d_grad, c->d_rdt, c->d_co,
                    c->d_ddtCoeff, c->d_Wold, c->d_Wold2, c->d_Wprev, c->d_src, c->d_dW, c->d_faceFlux, c->d_bad, c->d_offd, c->d_diag,
                    c->d_rD, c->d_invD, c->d_kry, c->d_w, c->d_x, c->d_scal, c->d_partial, c->d_counter, c->d_barrier, c->d_stage, c->d_lusgsYZ, c->d_lusgsHint, c->d_sliceRange,
                    c->d_faceRecon, c->d_gradE, c->d_visc, c->d_mrfFace, c->d_mrfOmega, c->d_transport, c->d_bfNbrPos, c->d_patchRot, c->d_bfAmiStart, c->d_amiAllSrc, c->d_amiAllW, c->d_rowLevF, c->d_rowLevR, c->d_tileNLevF, c->d_tileNLevR, c->d_tileDescF, c->d_tileDescR, c->d_blkTab, c->d_blkIdx, c->d_blkInfo, c->d_blkStage, c->d_blkFlag, c->d_blkCol, c->d_blkProf, c->d_blkTrace, c->d_hbD, c->d_hbPeer, c->d_hbInst, c->d_hbZone, c->d_hbZonePrm, c->d_hbInv, c->d_hbWork};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& pp : c->procs) if (pp.d_sendPos) cudaFree(pp.d_sendPos);
    for (auto& am : c->amis) { cudaFree(am.d_start); cudaFree(am.d_srcPos); cudaFree(am.d_w); }
    for (auto& ro : c->rots) { cudaFree(ro.d_srcPos); if (ro.d_lagSrc) cudaFree(ro.d_lagSrc); }
    for (auto& hb : c->h_bc) for (int fl = 0; fl < 3; fl++) if (hb.prmFace[fl]) cudaFree((void*)hb.prmFace[fl]);
    if (c->h_scal) cudaFreeHost(c->h_scal);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->evA);
    cudaEventDestroy(c->evB);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->commStream);
    delete c;
    return 0;
}

extern "C" const char* icsb200_last_error(icsb200_ctx* c) { return c ? c->err.c_str() : "null context"; }

extern "C" long long icsb200_launch_count(icsb200_ctx* c) { return c->launches; }

extern "C" int icsb200_timers_reset(icsb200_ctx* c, int enable)
{
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < TM_COUNT; i++) { c->tms[i] = 0; c->tcalls[i] = 0; }
    c->timing = enable != 0;
    return 0;
}

extern "C" int icsb200_timers_get(icsb200_ctx* c, char* names_buf, int names_len, double* ms, long long* calls, int max_classes)
{
    int n = std::min<int>(TM_COUNT, max_classes);
    int off = 0;
    for (int i = 0; i < n; i++) {
        int len = (int)std::strlen(kTimerNames[i]) + 1;
        if (names_buf && off + len <= names_len) { std::memcpy(names_buf + off, kTimerNames[i], len); off += len; }
        if (ms) ms[i] = c->tms[i];
        if (calls) calls[i] = c->tcalls[i];
    }
    return n;
}

extern "C" int icsb200_timer_begin(icsb200_ctx* c)
{
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    CUDA_TRY(c, cudaEventRecord(c->evA, c->stream));
    return 0;
}

extern "C" int icsb200_timer_end(icsb200_ctx* c, double* elapsed_ms)
{
    CUDA_TRY(c, cudaEventRecord(c->evB, c->stream));
    CUDA_TRY(c, cudaEventSynchronize(c->evB));
    float ms = 0;
    CUDA_TRY(c, cudaEventElapsedTime(&ms, c->evA, c->evB));
    *elapsed_ms = ms;
    return 0;
}

extern "C" int icsb200_schedule_info(icsb200_ctx* c, int out[8])
{
    out[0] = c->nLevF; out[1] = c->nLevR; out[2] = c->maxWidth; out[3] = c->NP;
    out[4] = c->tileMode ? 1 : 0; out[5] = c->nTiles; out[6] = c->nTileLevels; out[7] = c->blkMode ? 2 : (c->tileTma ? 1 : 0);
    return 0;
}

// ------------------------------------------------------------------------------------------------ selectors
extern "C" int icsb200_thermo_set(icsb200_ctx* c, double R, double Cp, double mu, double Pr)
{
    c->reconValid = false;
    if (!(R > 0) || !(Cp > R)) return ics_fail(c, ICSB200_EINVAL, "thermo: need Cp > R > 0");
    c->R = R; c->Cp = Cp; c->Cv = Cp - R; c->gamma = Cp / c->Cv; c->mu = mu; c->Pr = Pr;
    c->thermoSet = true;
    return 0;
}

extern "C" int icsb200_schemes_set(icsb200_ctx* c, const icsb200_schemes* s)
{
    c->reconValid = false;
    // newConvectiveFluxScheme.C:56-66: unknown type is a fatal error listing the valid ones
    if (s->flux_scheme < 0 || s->flux_scheme > ICSB200_FLUX_RUSANOV)
        return ics_fail(c, ICSB200_EINVAL, "Unknown convectiveFluxScheme type; valid types are: AUSMPlusUp HLLC ROE Rusanov");
    for (int l : {s->limiter_rho, s->limiter_U, s->limiter_T})
        if (l < 0 || l > ICSB200_LIM_LINEAR) return ics_fail(c, ICSB200_EINVAL, "unknown interpolation scheme for reconstruct(.)");
    if (s->ddt_scheme < 0 || s->ddt_scheme > ICSB200_DDT_BACKWARD) return ics_fail(c, ICSB200_EINVAL, "unknown ddt scheme");
    if (s->ddt_scheme != ICSB200_DDT_STEADY && !(s->delta_t > 0)) return ics_fail(c, ICSB200_EINVAL, "transient run needs delta_t > 0");
    c->sch = *s;
    c->pseudoCoNum = s->pseudo_co_num;
    return 0;
}

extern "C" int icsb200_bc_set(icsb200_ctx* c, int patch, int field, int kind, const double* params, int n_params)
{
    c->reconValid = false;
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "bc_set: mesh not set");
    if (patch < 0 || patch >= (int)c->patches.size() || field < 0 || field > 2 || n_params < 0 || n_params > 8 || kind < 0 ||
        kind > ICSB200_BC_COUPLED)
        return ics_fail(c, ICSB200_EINVAL, "bc_set: bad argument");
    c->h_bc[patch].kind[field] = kind;
    for (int i = 0; i < 8; i++) c->h_bc[patch].prm[field][i] = i < n_params ? params[i] : 0.0;
    if (c->h_bc[patch].prmFace[field]) { cudaFree((void*)c->h_bc[patch].prmFace[field]); c->h_bc[patch].prmFace[field] = nullptr; }
    CUDA_TRY(c, cudaMemcpy(c->d_bc, c->h_bc.data(), sizeof(BCDev) * c->h_bc.size(), cudaMemcpyHostToDevice));
    return 0;
}

// non-uniform patch-field entries (`nonuniform List<...>` of value / p0 / T0 / inletValue / tangentialVelocity in 0/p, 0/U, 0/T):
// params[size of the patch][n_params], one row per face in patch face order
extern "C" int icsb200_bc_set_nonuniform(icsb200_ctx* c, int patch, int field, int kind, const double* params, int n_params)
{
    c->reconValid = false;
    if (!c->meshSet) return ics_fail(c, ICSB200_ESTATE, "bc_set: mesh not set");
    if (patch < 0 || patch >= (int)c->patches.size() || field < 0 || field > 2 || n_params < 1 || n_params > 8 || kind < 0 ||
        kind > ICSB200_BC_COUPLED || !params)
        return ics_fail(c, ICSB200_EINVAL, "bc_set: bad argument");
    cudaSetDevice(c->device);
    const int n = c->patches[patch].size;
    std::vector<double> h((size_t)8 * std::max(n, 1), 0.0);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n_params; k++) h[(size_t)8 * i + k] = params[(size_t)n_params * i + k];
    double* d = const_cast<double*>(c->h_bc[patch].prmFace[field]);
    int r = devUpload(c, &d, h);
    if (r) return r;
    c->h_bc[patch].kind[field] = kind;
    c->h_bc[patch].prmFace[field] = d;
    c->h_bc[patch].bstart = c->patches[patch].start - c->F;
    CUDA_TRY(c, cudaMemcpy(c->d_bc, c->h_bc.data(), sizeof(BCDev) * c->h_bc.size(), cudaMemcpyHostToDevice));
    return 0;
}

// ------------------------------------------------------------------------------------------------ mesh
extern "C" int icsb200_ami_set(icsb200_ctx* c, int patch, int n_faces, const int* face_start, const int* nbr_face, const double* weight)
{
    if (patch < 0 || n_faces < 0 || !face_start || (face_start[n_faces] > 0 && (!nbr_face || !weight)))
        return ics_fail(c, ICSB200_EINVAL, "ami_set: bad arguments");
    AmiTable t;
    t.start.assign(face_start, face_start + n_faces + 1);
    t.face.assign(nbr_face, nbr_face + face_start[n_faces]);
    t.weight.assign(weight, weight + face_start[n_faces]);
    for (auto& pa : c->pendingAmi) if (pa.first == patch) { pa.second = t; return 0; }
    c->pendingAmi.emplace_back(patch, t);
    return 0;
}

extern "C" int icsb200_phaselag_set(icsb200_ctx* c, int patch, int n_instants, const double* weights)
{
    if (patch < 0 || n_instants < 2 || n_instants > 16 || !weights) return ics_fail(c, ICSB200_EINVAL, "phaselag_set: bad arguments");
    std::vector<double> w(weights, weights + n_instants);
    for (auto& pe : c->pendingLag) if (pe.first == patch) { pe.second = w; return 0; }
    c->pendingLag.emplace_back(patch, w);
    return 0;
}

extern "C" int icsb200_mesh_set(icsb200_ctx* c, int N, int F, int FT, const int* owner, const int* neighbour, const double* Sf,
                                const double* magSf, const double* weights, const double* deltaCoeffs, const double* nonOrthDeltaCoeffs,
                                const double* C, const double* V, const double* Cf, int n_patches, const icsb200_patch* patches,
                                const int solutionD[3])
{
    if (N <= 0 || F < 0 || FT < F) return ics_fail(c, ICSB200_EINVAL, "mesh_set: bad sizes");
    cudaSetDevice(c->device);
    for (int f = 0; f < F; f++)
        if (owner[f] >= neighbour[f] || owner[f] < 0 || neighbour[f] >= N)
            return ics_fail(c, ICSB200_EINVAL, "mesh_set: faces must be in upper-triangular order (owner < neighbour)");
    c->N = N; c->F = F; c->FT = FT; c->NB = FT - F;
    c->owner.assign(owner, owner + FT);
    c->neighbour.assign(neighbour, neighbour + F);
    c->patches.assign(patches, patches + n_patches);
    for (int d = 0; d < 3; d++) c->solutionD[d] = solutionD[d];
    const int NB = c->NB;
    c->bfacePatch.assign(NB, -1);
    auto lagOf = [&](int pi) -> const std::vector<double>* {
        for (auto& pe : c->pendingLag) if (pe.first == pi) return &pe.second;
        return nullptr;
    };
    // cyclic patches whose neighbour values are not a plain copy live in local halo slots: rotational pairs and phase-lag pairs
    auto localHalo = [&](int pi) { return patches[pi].kind == ICSB200_CYCLIC && (ics_is_rotational(patches[pi]) || lagOf(pi)); };
    for (auto& pe : c->pendingLag) {
        if (pe.first >= n_patches || patches[pe.first].kind != ICSB200_CYCLIC) return ics_fail(c, ICSB200_EINVAL, "mesh_set: phaselag_set on a patch that is not cyclic");
        if (N % (int)pe.second.size()) return ics_fail(c, ICSB200_EINVAL, "mesh_set: phase-lag patch on a mesh that is not made of n_instants copies");
    }
    for (int pi = 0; pi < n_patches; pi++) {
        const icsb200_patch& p = patches[pi];
        if (p.start < F || p.start + p.size > FT) return ics_fail(c, ICSB200_EINVAL, "mesh_set: patch range outside boundary faces");
        if ((p.kind == ICSB200_CYCLIC) && (p.nbr_patch < 0 || p.nbr_patch >= n_patches || patches[p.nbr_patch].size != p.size))
            return ics_fail(c, ICSB200_EINVAL, "mesh_set: cyclic patch without a matching neighbour patch");
        if (p.kind == ICSB200_CYCLICAMI) {
            if (p.nbr_patch < 0 || p.nbr_patch >= n_patches || patches[p.nbr_patch].kind != ICSB200_CYCLICAMI)
                return ics_fail(c, ICSB200_EINVAL, "mesh_set: cyclicAMI patch without a cyclicAMI neighbour patch");
            const AmiTable* t = nullptr;
            for (auto& pa : c->pendingAmi) if (pa.first == pi) t = &pa.second;
            if (!t || (int)t->start.size() != p.size + 1) return ics_fail(c, ICSB200_EINVAL, "mesh_set: cyclicAMI patch without a matching icsb200_ami_set");
            for (int fidx : t->face) if (fidx < 0 || fidx >= patches[p.nbr_patch].size) return ics_fail(c, ICSB200_EINVAL, "mesh_set: cyclicAMI address outside the neighbour patch");
        }
        // rotational pairs (forwardT != I): plain cyclic only; patchNeighbourField = transform(forwardT, neighbour value)
        // (originalOFFiles/constraintFvPatchFields/cyclic/cyclicFvPatchField.C:130-190) through local halo slots

        if (p.kind == ICSB200_PROCESSOR && (p.nbr_rank < 0 || p.nbr_rank >= c->nRanks || c->nRanks == 1))
            return ics_fail(c, ICSB200_EINVAL, "mesh_set: processor patch needs a multi-rank context");
        if ((p.kind == ICSB200_CYCLIC || p.kind == ICSB200_PROCESSOR || p.kind == ICSB200_CYCLICAMI) && !Cf) return ics_fail(c, ICSB200_EINVAL, "mesh_set: coupled patches need Cf");
        if (p.kind != ICSB200_EMPTY) for (int f = p.start; f < p.start + p.size; f++) c->bfacePatch[f - F] = pi;
    }

    // ---- LU-SGS schedule.  Levels: forward = longest path from below, reverse = longest path from above.
    std::vector<int> levF(N, 0), levR(N, 0);
    for (int f = 0; f < F; f++) levF[neighbour[f]] = std::max(levF[neighbour[f]], levF[owner[f]] + 1);
    for (int f = F - 1; f >= 0; f--) levR[owner[f]] = std::max(levR[owner[f]], levR[neighbour[f]] + 1);
    int nLevF = 0, nLevR = 0;
    for (int i = 0; i < N; i++) { nLevF = std::max(nLevF, levF[i] + 1); nLevR = std::max(nLevR, levR[i] + 1); }
    c->nLevF = nLevF; c->nLevR = nLevR;
    c->maxWidth = 0;
    {
        std::vector<int> cnt(nLevF, 0);
        for (int i = 0; i < N; i++) c->maxWidth = std::max(c->maxWidth, ++cnt[levF[i]]);
    }
    // Tile mode (blocked wavefront): cells are binned into boxes of ~512 cells by coordinate thresholds.  If every
    // internal face goes from a tile to a component-wise >= tile, the tile graph is acyclic with tile level a+b+c and
    // ordering positions by (tile level, tile, forward level, cell) is a valid topological order in which a CTA can sweep
    // a whole tile with shared memory, synchronising with other CTAs only once per tile.  Otherwise fall back to the
    // level order (one 32-row slice per "tile").
    std::vector<int> tileOf;  // per cell
    std::vector<int> tileCol; // block tiles: column of a tile (tiles of a column are consecutive, chunks ascending) or empty
    bool colMode = false;
    int nTiles = 0;
    c->tileMode = false;
    bool blkWanted = false;
    {
        const char* env = getenv("ICSB200_LUSGS_MODE");
        const std::string mode = env ? std::string(env) : std::string("auto");
        // "auto" / "blk": ~512-row block tiles swept by k_lusgs_blk (lusgs_blk.cu) — the default whenever the mesh allows it;
        // "level": the level pipeline (k_lusgs_tma); "tile64": 64-row tiles for the TMA tile kernel; "tile": the older ~512-row tiles
        const bool want64 = mode == "tile64";
        bool wantBlk = mode == "auto" || mode == "blk";
        if (wantBlk) {
            // the block-tile kernel stages at most 3 lower and 3 upper neighbours per row (hex-like cells)
            std::vector<unsigned char> nl(N, 0), nu(N, 0);
            for (int f = 0; f < F && wantBlk; f++)
                if (++nl[neighbour[f]] > 3 || ++nu[owner[f]] > 3) wantBlk = false;
        }
        const bool wantTiles = (mode == "tile" || want64 || wantBlk) && N >= 64;
        const double tileTarget = want64 ? 64.0 : (wantBlk ? (double)ICS_BLK_MR : 512.0);
        const int tileRowCap = want64 ? 64 : (wantBlk ? ICS_BLK_MR : ICS_TILE_MAXROWS - 32);
        c->tileTma = false;
        if (wantTiles) {
            // logical coordinates from the graph alone: u_d = longest path using only faces whose normal is mostly along
            // axis d (exactly (i,j,k) on a block-structured mesh, however curved); tiles = boxes of u
            std::vector<int> u[3];
            int umax[3] = {0, 0, 0};
            for (int d = 0; d < 3; d++) u[d].assign(N, 0);
            // Block tiles: every face constrains all three coordinates (weight 1 along its own axis, 0 along the others), forwards
            // (u(neighbour) >= u(owner) + w) and backwards (u(owner) >= u(neighbour) - w), relaxed until nothing moves: on a
            // logically structured mesh that is the exact index potential (i,j,k) up to a shift, whatever the shape of the
            // sub-domain.  A forward longest path alone starts every row of a staircase-shaped partition cut (x-slabs of a curved
            // mesh) at zero, which put several cells on one coordinate and made a y-face step back in x: three of the eight
            // bump-4M slabs fell back to the level pipeline.  The last pass is a forward one, so every coordinate is
            // non-decreasing across every face even where the mesh is not structured (the tile graph is checked anyway).
            const bool legacyU = !wantBlk;   // the older tile modes keep their coordinates
            std::vector<unsigned char> axisOf(F);
            for (int f = 0; f < F; f++) {
                double ax = std::fabs(Sf[3 * (size_t)f]), ay = std::fabs(Sf[3 * (size_t)f + 1]), az = std::fabs(Sf[3 * (size_t)f + 2]);
                axisOf[f] = (unsigned char)((ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2));
            }
            auto forwardPass = [&]() {
                bool changed = false;
                for (int f = 0; f < F; f++) {
                    const int d = axisOf[f], o = owner[f], n2 = neighbour[f];
                    for (int a2 = 0; a2 < 3; a2++) {
                        if (legacyU && a2 != d) continue;
                        const int v = u[a2][o] + (a2 == d ? 1 : 0);
                        if (v > u[a2][n2]) { u[a2][n2] = v; changed = true; }
                    }
                }
                return changed;
            };
            auto backwardPass = [&]() {
                bool changed = false;
                for (int f = F - 1; f >= 0; f--) {
                    const int d = axisOf[f], o = owner[f], n2 = neighbour[f];
                    for (int a2 = 0; a2 < 3; a2++) {
                        const int v = u[a2][n2] - (a2 == d ? 1 : 0);
                        if (v > u[a2][o]) { u[a2][o] = v; changed = true; }
                    }
                }
                return changed;
            };
            forwardPass();
            if (!legacyU)
                for (int it = 0; it < 4; it++) {
                    if (!backwardPass()) break;
                    if (!forwardPass()) break;
                }
            for (int d = 0; d < 3; d++)
                for (int i = 0; i < N; i++) umax[d] = std::max(umax[d], u[d][i]);
            int active = 0;
            for (int d = 0; d < 3; d++) if (umax[d] + 1 >= 4) active++;
            int side = active > 0 ? std::max(2, (int)std::lround(std::pow(tileTarget, 1.0 / active))) : (int)tileTarget;
            if (wantBlk && active > 0) side = std::max(2, (int)std::floor(std::pow(tileTarget, 1.0 / active) + 1e-9));  // side^active <= cap
            // disconnected copies of a mesh (Harmonic Balance instances) share logical coordinates: the connected component
            // is a fourth tile coordinate (block tiles only; the older modes keep their layout)
            std::vector<int> comp;
            int nComp = 1;
            if (wantBlk) {
                std::vector<int> parent(N);
                for (int i = 0; i < N; i++) parent[i] = i;
                auto find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
                for (int f = 0; f < F; f++) { int a = find(owner[f]), b = find(neighbour[f]); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
                comp.assign(N, -1);
                nComp = 0;
                for (int i = 0; i < N; i++) { int r0 = find(i); if (comp[r0] < 0) comp[r0] = nComp++; comp[i] = comp[r0]; }
            }
            if (wantBlk) {
                // ---- skewed column tiles.  The longest active axis is the sweep axis; the other active axes are cut into
                // blocks of bx (x by) cells, which gives columns along the sweep axis.  A cell's local level is
                // s = sum over the cross axes of (u mod b) + u_sweep — the forward hyperplane index inside its column — and a
                // tile is `depth` consecutive local levels of one column.  Away from the column ends every intra-tile level is a
                // full bx*by rows wide (a cube tile of the same size spends most of its levels filling and draining the
                // wavefront: 22 levels of 23 rows on average for 8^3 against 8 levels of 64 rows), and the dependency chain
                // between tiles stays the hyperplane count / depth.  A face that leaves a column through a block boundary goes
                // from local level s to s - (b-1), i.e. at most ceil((b-1)/depth) chunks back, so TL = sum K_a bin_a + chunk with
                // K_a = ceil((b_a-1)/depth) + 1 ranks the tile graph; every face is checked against it (any mesh that fails
                // falls back to the level pipeline).
                int sweep = -1;
                for (int d = 0; d < 3; d++) if (umax[d] + 1 >= 4 && (sweep < 0 || umax[d] > umax[sweep])) sweep = d;
                auto envInt = [](const char* name, int dflt) { const char* e = getenv(name); return e ? std::max(1, atoi(e)) : dflt; };
                const int nCross = active > 0 ? active - 1 : 0;
                // 3-D: 8 x 4 columns (32 rows = one warp per block-row component and level), 8 or 4 levels deep; 2-D: 32 wide, 8 deep
                // (profiles/r02_lusgs_tileshape_sweep.log, r02_lusgs_colshape_sweep.log)
                int bside = nCross == 2 ? 8 : (nCross == 1 ? 32 : 1);
                bside = envInt("ICSB200_LUSGS_BSIDE", bside);
                int bside2 = nCross == 2 ? 4 : bside;
                bside2 = envInt("ICSB200_LUSGS_BSIDE2", bside2);
                // depth (3-D): 8 levels per tile amortise the per-tile traffic and hop best when the sweep is bandwidth-bound (344^3:
                // 10.3 ms against 12.8 with 4); when it is bound by the chain of tile steps, the depth with the shorter chain wins
                // (172^3: 193 steps of 4 levels, 2.68 ms, against 149 steps of 8 levels, 2.82 ms).  Model: step = depth x 0.6 us +
                // 4.6 us (lusgs_blk_trace.py medians), bytes at 5.5 TB/s.
                int depth = nCross == 2 ? 8 : (nCross == 1 ? 8 : 32);
                // COLUMN mode (default wherever there are columns): a CTA sweeps all chunks of a column back to back and hands each
                // chunk's last level to the next chunk in shared memory, so only the lateral dependencies (neighbour columns) cross
                // L2: 172^3 2.73 -> 1.95 ms, 344^3 10.35 -> 9.70 ms, bump-4M 7.2 -> 4.9 ms.
                colMode = nCross >= 1;
                if (nCross == 2) {
                    auto steps = [&](int dep) {
                        int n = 0, sm2 = umax[sweep], nth = 0;
                        for (int d = 0; d < 3; d++) {
                            if (d == sweep) continue;
                            if (umax[d] + 1 < 4) { sm2 += umax[d]; continue; }
                            const int b2 = nth++ == 0 ? bside : bside2;
                            n += ((b2 - 1 + dep - 1) / dep + 1) * (umax[d] / b2);
                            sm2 += b2 - 1;
                        }
                        return n + sm2 / dep + 1;
                    };
                    const double chain8 = steps(8) * (8 * 0.6 + 4.6), chain4 = steps(4) * (4 * 0.6 + 4.6);   // us per sweep
                    const double bytesUs = (double)N * 710.0 / 5.5e6;   // us per sweep at 5.5 TB/s
                    depth = (bytesUs < chain8 && chain4 < chain8) ? 4 : 8;
                    // chain-bound meshes in column mode: the chain is the sum over the cross axes of bins x (K T + 4.6 us) + chunks x T
                    // with T = depth x 0.6 us
                    if (bytesUs < chain8) {
                        auto colChain = [&](int dep) {
                            double tsum = 0.0;
                            int sm2 = umax[sweep], nth = 0;
                            const double T = dep * 0.6;
                            for (int d = 0; d < 3; d++) {
                                if (d == sweep) continue;
                                if (umax[d] + 1 < 4) { sm2 += umax[d]; continue; }
                                const int b2 = nth++ == 0 ? bside : bside2;
                                tsum += (umax[d] / b2) * (((b2 - 1 + dep - 1) / dep + 1) * T + 4.6);
                                sm2 += b2 - 1;
                            }
                            return tsum + (sm2 / dep + 1) * T;
                        };
                        // measured: 172^3 1.95 ms (4) against 2.00 (8), 200^3 2.78 (4) against 2.58 (8), 344^3 12.5 (4) against 9.7 (8):
                        // the shallow tiles only pay on small partitions
                        depth = (N < 6000000 && colChain(4) < colChain(8)) ? 4 : 8;
                    }
                }
                if (const char* e = getenv("ICSB200_LUSGS_COLMODE")) colMode = atoi(e) != 0;
                depth = envInt("ICSB200_LUSGS_DEPTH", depth);
                int bs[3] = {1, 1, 1}, nbin[3] = {1, 1, 1}, K[3] = {0, 0, 0};
                {
                    int nth = 0;
                    for (int d = 0; d < 3; d++) {
                        if (d == sweep || umax[d] + 1 < 4) continue;
                        bs[d] = nth++ == 0 ? bside : bside2;
                        nbin[d] = umax[d] / bs[d] + 1;
                        K[d] = (bs[d] - 1 + depth - 1) / depth + 1;
                    }
                }
                int sMax = sweep >= 0 ? umax[sweep] : 0;
                for (int d = 0; d < 3; d++) if (d != sweep) sMax += (umax[d] + 1 < 4) ? umax[d] : bs[d] - 1;
                const int nChunk = sMax / depth + 1;
                const long long nb3 = (long long)nbin[0] * nbin[1] * nbin[2] * nChunk;
                const long long nbAll = nb3 * nComp;
                bool ok = sweep >= 0 && nbAll <= (1ll << 30) && (long long)bs[0] * bs[1] * bs[2] * depth <= tileRowCap;
                if (ok) {
                    std::vector<int> key(N), tlOf(N);
                    for (int i = 0; i < N; i++) {
                        int s2 = u[sweep][i], b3 = 0, mul = 1, tl2 = 0;
                        for (int d = 0; d < 3; d++) {
                            if (d == sweep) continue;
                            // inactive axes (fewer than 4 layers) keep their full extent inside the column
                            if (bs[d] == 1 && nbin[d] == 1) { s2 += u[d][i]; continue; }
                            const int bn = u[d][i] / bs[d];
                            s2 += u[d][i] - bn * bs[d];
                            b3 += mul * bn; mul *= nbin[d];
                            tl2 += K[d] * bn;
                        }
                        const int ch = std::min(s2 / depth, nChunk - 1);
                        key[i] = b3 + mul * ch + (nComp > 1 ? (int)(nb3 * comp[i]) : 0);
                        tlOf[i] = tl2 + ch;
                    }
                    for (int f = 0; f < F && ok; f++)
                        if (key[owner[f]] != key[neighbour[f]] && tlOf[owner[f]] >= tlOf[neighbour[f]]) {
                            ok = false;
                            if (getenv("ICSB200_LUSGS_DEBUG")) {
                                const int o = owner[f], n2 = neighbour[f];
                                fprintf(stderr, "icsb200: tile graph not ranked at face %d: owner %d u=(%d,%d,%d) TL %d -> neighbour %d u=(%d,%d,%d) TL %d (sweep axis %d, blocks %d %d %d, depth %d)\n",
                                        f, o, u[0][o], u[1][o], u[2][o], tlOf[o], n2, u[0][n2], u[1][n2], u[2][n2], tlOf[n2], sweep, bs[0], bs[1], bs[2], depth);
                            }
                        }
                    if (ok) {
                        // tile order: by tile level; inside a level by descending chunk = ascending column index sum, so that every
                        // tile a tile waits for sits at the same or an earlier relative place of the previous level — with
                        // round-robin dealing of tiles to CTAs that maximises the distance (in tiles) to a dependency
                        std::vector<int> cntT(nbAll, 0), tlT(nbAll, 0), chT(nbAll, 0);
                        for (int i = 0; i < N; i++) { cntT[key[i]]++; tlT[key[i]] = tlOf[i]; }
                        {
                            long long mulAll = 1;
                            for (int d = 0; d < 3; d++) if (d != sweep && !(bs[d] == 1 && nbin[d] == 1)) mulAll *= nbin[d];
                            for (long long t = 0; t < nbAll; t++) chT[t] = (int)((t % nb3) / mulAll);
                        }
                        std::vector<int> order;
                        for (long long t = 0; t < nbAll; t++) if (cntT[t] > 0) order.push_back((int)t);
                        // column of a tile: its key without the chunk; column mode orders the tiles by (column rank, chunk), columns
                        // ranked by their K-weighted bin sum (every neighbour column a column waits for has a smaller one)
                        std::vector<long long> colT(nbAll, 0);
                        {
                            long long mulAll = 1;
                            for (int d = 0; d < 3; d++) if (d != sweep && !(bs[d] == 1 && nbin[d] == 1)) mulAll *= nbin[d];
                            for (long long t = 0; t < nbAll; t++) colT[t] = (t % nb3) % mulAll + (t / nb3) * mulAll;
                        }
                        if (colMode)
                            std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
                                const int ax = tlT[x] - chT[x], ay = tlT[y] - chT[y];
                                if (ax != ay) return ax < ay;
                                if (colT[x] != colT[y]) return colT[x] < colT[y];
                                return chT[x] < chT[y];
                            });
                        else
                            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tlT[x] != tlT[y] ? tlT[x] < tlT[y] : chT[x] > chT[y]; });
                        int maxRows = 0, maxTL = 0;
                        for (int t : order) { maxRows = std::max(maxRows, cntT[t]); maxTL = std::max(maxTL, tlT[t]); }
                        if (maxRows > tileRowCap && getenv("ICSB200_LUSGS_DEBUG")) fprintf(stderr, "icsb200: a tile has %d rows (cap %d)\n", maxRows, tileRowCap);
                        if (maxRows <= tileRowCap) {
                            std::vector<int> dense(nbAll, -1);
                            for (size_t k = 0; k < order.size(); k++) dense[order[k]] = (int)k;
                            tileOf.resize(N);
                            for (int i = 0; i < N; i++) tileOf[i] = dense[key[i]];
                            nTiles = (int)order.size();
                            if (colMode) {
                                tileCol.resize(nTiles);
                                int nc = -1;
                                for (int k = 0; k < nTiles; k++) {
                                    if (k == 0 || colT[order[k]] != colT[order[k - 1]]) nc++;
                                    tileCol[k] = nc;
                                }
                            }
                            c->tileMode = true;
                            blkWanted = true;
                            c->nTileLevels = maxTL + 1;
                        }
                    }
                }
            } else {
                int nb[3];
                std::vector<int> bin[3];
                for (int d = 0; d < 3; d++) {
                    const bool act = umax[d] + 1 >= 4;
                    nb[d] = act ? (umax[d] + side) / side : 1;
                    bin[d].assign(N, 0);
                    if (act) for (int i = 0; i < N; i++) bin[d][i] = u[d][i] / side;
                }
                bool ok = true;
                for (int f = 0; f < F && ok; f++)
                    for (int d = 0; d < 3; d++)
                        if (bin[d][neighbour[f]] < bin[d][owner[f]]) { ok = false; break; }
                const long long nb3 = (long long)nb[0] * nb[1] * nb[2];
                const long long nbAll = nb3 * nComp;
                if (nbAll > (1ll << 30)) ok = false;
                if (ok) {
                    // dense tile numbering sorted by (tile level, component, a, b, c)
                    std::vector<int> key(N);
                    std::vector<int> cntT(nbAll, 0);
                    for (int i = 0; i < N; i++) {
                        key[i] = bin[0][i] + nb[0] * (bin[1][i] + nb[1] * bin[2][i]) + (nComp > 1 ? (int)(nb3 * comp[i]) : 0);
                        cntT[key[i]]++;
                    }
                    std::vector<int> order;
                    for (long long t = 0; t < nbAll; t++) if (cntT[t] > 0) order.push_back((int)t);
                    auto tl = [&](int t) { const int t3 = (int)(t % nb3); return t3 % nb[0] + (t3 / nb[0]) % nb[1] + t3 / (nb[0] * nb[1]); };
                    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return tl(x) < tl(y); });
                    int maxRows = 0, maxTL = 0;
                    for (int t : order) { maxRows = std::max(maxRows, cntT[t]); maxTL = std::max(maxTL, tl(t)); }
                    if (maxRows <= tileRowCap) {
                        std::vector<int> dense(nbAll, -1);
                        for (size_t k = 0; k < order.size(); k++) dense[order[k]] = (int)k;
                        tileOf.resize(N);
                        for (int i = 0; i < N; i++) tileOf[i] = dense[key[i]];
                        nTiles = (int)order.size();
                        c->tileMode = true;
                        c->tileTma = want64;
                        blkWanted = wantBlk;
                        c->nTileLevels = maxTL + 1;
                    }
                }
            }
        }
    }
    std::vector<int> tileStart;                 // [nTiles+1] first position of a tile (multiple of 32)
    std::vector<int> tileFPtr, tileFLev;        // forward: CSR over tiles of local row offsets where a new level starts
    std::vector<int> tileRPtr, tileRLev, tileRRows;  // reverse: rows of a tile sorted by reverse level + level offsets
    int NP = 0;
    if (c->tileMode) {
        std::vector<int> cntT(nTiles, 0);
        for (int i = 0; i < N; i++) cntT[tileOf[i]]++;
        tileStart.assign(nTiles + 1, 0);
        for (int t = 0; t < nTiles; t++) tileStart[t + 1] = tileStart[t] + ((cntT[t] + 31) / 32) * 32;
        NP = tileStart[nTiles];
        c->pos2cell.assign(NP, -1);
        c->cell2pos.assign(N, -1);
        // cells of a tile sorted by (forward level, cell id): bucket by tile, then sort each bucket
        std::vector<int> fill(tileStart.begin(), tileStart.end() - 1);
        for (int i = 0; i < N; i++) c->pos2cell[fill[tileOf[i]]++] = i;
        tileFPtr.assign(nTiles + 1, 0); tileRPtr.assign(nTiles + 1, 0);
        tileRRows.assign(NP, 0);
        for (int t = 0; t < nTiles; t++) {
            int* beg = c->pos2cell.data() + tileStart[t];
            int n = cntT[t];
            std::sort(beg, beg + n, [&](int x, int y) { return levF[x] != levF[y] ? levF[x] < levF[y] : x < y; });
            for (int r = 0; r < n; r++) {
                c->cell2pos[beg[r]] = tileStart[t] + r;
                if (r == 0 || levF[beg[r]] != levF[beg[r - 1]]) tileFLev.push_back(r);
            }
            tileFLev.push_back(n);
            tileFPtr[t + 1] = (int)tileFLev.size();
            // reverse order: local rows sorted by (reverse level, local row)
            std::vector<int> rr(n);
            for (int r = 0; r < n; r++) rr[r] = r;
            std::sort(rr.begin(), rr.end(), [&](int x, int y) { return levR[beg[x]] != levR[beg[y]] ? levR[beg[x]] < levR[beg[y]] : x < y; });
            for (int r = 0; r < n; r++) {
                tileRRows[tileStart[t] + r] = rr[r];
                if (r == 0 || levR[beg[rr[r]]] != levR[beg[rr[r - 1]]]) tileRLev.push_back(r);
            }
            tileRLev.push_back(n);
            tileRPtr[t + 1] = (int)tileRLev.size();
        }
        c->nTiles = nTiles;
    } else {
        // level order: positions sorted by (forward level, cell id), levels padded to a multiple of 32
        std::vector<int> cntF(nLevF + 1, 0);
        for (int i = 0; i < N; i++) cntF[levF[i] + 1]++;
        std::vector<int> levStartF(nLevF + 1, 0);
        for (int l = 0; l < nLevF; l++) levStartF[l + 1] = levStartF[l] + ((cntF[l + 1] + 31) / 32) * 32;
        NP = levStartF[nLevF];
        c->pos2cell.assign(NP, -1);
        c->cell2pos.assign(N, -1);
        std::vector<int> fill(levStartF.begin(), levStartF.end() - 1);
        for (int i = 0; i < N; i++) { int p = fill[levF[i]]++; c->pos2cell[p] = i; c->cell2pos[i] = p; }
        c->nTiles = 0;
        c->nTileLevels = 0;
    }
    c->NP = NP;

    // ---- halo slots for processor patches
    c->procs.clear();
    int NH = 0;
    std::vector<int> patchHaloStart(n_patches, -1);
    for (int pi = 0; pi < n_patches; pi++)
        if (patches[pi].kind == ICSB200_PROCESSOR || patches[pi].kind == ICSB200_CYCLICAMI || localHalo(pi)) { patchHaloStart[pi] = NH; NH += patches[pi].size; }
    NH += NH & 1;  // keep NPH even: every component array of a cell vector then starts 16-byte aligned (TMA bulk copies)
    c->NH = NH; c->NPH = NP + NH; c->NX = NP + NH + NB;

    // ---- rows: faces of each cell in ascending face id
    std::vector<int> nLowC(N, 0), nUpC(N, 0), nBndC(N, 0);
    for (int f = 0; f < F; f++) { nLowC[neighbour[f]]++; nUpC[owner[f]]++; }
    for (int b = 0; b < NB; b++) if (c->bfacePatch[b] >= 0) nBndC[owner[F + b]]++;
    const int nSlices = NP / 32;
    c->nSlices = nSlices;
    c->h_sliceOff.assign(nSlices + 1, 0);
    c->h_rowNLow.assign(NP, 0); c->h_rowNInt.assign(NP, 0); c->h_rowNAll.assign(NP, 0);
    for (int s = 0; s < nSlices; s++) {
        int w = 0;
        for (int l = 0; l < 32; l++) {
            int p = s * 32 + l, i = c->pos2cell[p];
            if (i < 0) continue;
            c->h_rowNLow[p] = nLowC[i];
            c->h_rowNInt[p] = nLowC[i] + nUpC[i];
            c->h_rowNAll[p] = nLowC[i] + nUpC[i] + nBndC[i];
            w = std::max(w, c->h_rowNAll[p]);
        }
        c->h_sliceOff[s + 1] = c->h_sliceOff[s] + w;
    }
    c->nEntries = c->h_sliceOff[nSlices];
    const size_t nE32 = (size_t)c->nEntries * 32;
    if (nE32 * 25 > (size_t)1 << 40) return ics_fail(c, ICSB200_EINVAL, "mesh too large");
    c->h_col.assign(nE32, -1); c->h_meta.assign(nE32, ET_PHYS); c->h_gfid.assign(nE32, -1);
    {
        std::vector<int> fillLow(N, 0), fillUp(N, 0), fillB(N, 0);
        auto slot = [&](int p, int j) { return ((size_t)c->h_sliceOff[p / 32] + j) * 32 + (p % 32); };
        for (int f = 0; f < F; f++) {
            int o = owner[f], n = neighbour[f];
            int pn = c->cell2pos[n], po = c->cell2pos[o];
            size_t sl = slot(pn, fillLow[n]++);
            c->h_col[sl] = po; c->h_meta[sl] = ET_LOWER | (f << 2);
            size_t su = slot(po, nLowC[o] + fillUp[o]++);
            c->h_col[su] = pn; c->h_meta[su] = ET_UPPER | (f << 2);
        }
        c->bfNbrSlot.assign(NB, -1);
        for (int b = 0; b < NB; b++) {
            int pi = c->bfacePatch[b];
            if (pi < 0) continue;
            int f = F + b, o = owner[f], po = c->cell2pos[o];
            size_t sb = slot(po, nLowC[o] + nUpC[o] + fillB[o]++);
            const icsb200_patch& pa = patches[pi];
            if (pa.kind == ICSB200_CYCLIC && !localHalo(pi)) {
                int nbrFace = patches[pa.nbr_patch].start + (f - pa.start);
                c->h_col[sb] = c->cell2pos[owner[nbrFace]];
                c->h_meta[sb] = ET_COUPLED | (f << 2);
            } else if (pa.kind == ICSB200_CYCLIC) {   // rotational / phase lag: the neighbour values live in a local halo slot
                c->h_col[sb] = NP + patchHaloStart[pi] + (f - pa.start);
                c->h_meta[sb] = ET_COUPLED | (f << 2);
            } else if (pa.kind == ICSB200_PROCESSOR || pa.kind == ICSB200_CYCLICAMI) {
                c->h_col[sb] = NP + patchHaloStart[pi] + (f - pa.start);
                c->h_meta[sb] = ET_COUPLED | (f << 2);
            } else {
                c->h_col[sb] = NP + NH + b;  // boundary-value slot
                c->h_meta[sb] = ET_PHYS | (f << 2);
            }
            if ((c->h_meta[sb] & 3) == ET_COUPLED) c->bfNbrSlot[b] = c->h_col[sb];
        }
    }
    // per-slice entry ranges the LU-SGS sweeps touch: forward [0, max nLow), reverse [min nLow over rows with uppers, max nInt)
    {
        std::vector<int> range((size_t)3 * nSlices, 0);
        for (int s = 0; s < nSlices; s++) {
            int fHi = 0, rLo = 1 << 20, rHi = 0;
            for (int l = 0; l < 32; l++) {
                const int p = s * 32 + l;
                fHi = std::max(fHi, c->h_rowNLow[p]);
                if (c->h_rowNInt[p] > c->h_rowNLow[p]) { rLo = std::min(rLo, c->h_rowNLow[p]); rHi = std::max(rHi, c->h_rowNInt[p]); }
            }
            if (rHi == 0) rLo = 0;
            range[s] = fHi; range[(size_t)nSlices + s] = rLo; range[2 * (size_t)nSlices + s] = rHi;
        }
        int rr = devUpload(c, &c->d_sliceRange, range);
        if (rr) return rr;
    }
    // GPU face ids: owner-side entries in (slice, j, lane) order
    std::vector<int> ref2gf(FT, -1);
    c->h_gf2ref.clear();
    c->h_gf2ref.reserve((size_t)F + NB);
    for (int s = 0; s < nSlices; s++) {
        int w = c->h_sliceOff[s + 1] - c->h_sliceOff[s];
        for (int j = 0; j < w; j++)
            for (int l = 0; l < 32; l++) {
                int p = s * 32 + l;
                if (j >= c->h_rowNAll[p] || j < c->h_rowNLow[p]) continue;
                size_t sl = ((size_t)c->h_sliceOff[s] + j) * 32 + l;
                int f = c->h_meta[sl] >> 2;
                ref2gf[f] = (int)c->h_gf2ref.size();
                c->h_gf2ref.push_back(f);
            }
    }
    c->NFG = (int)c->h_gf2ref.size();
    for (int s = 0; s < nSlices; s++) {
        int w = c->h_sliceOff[s + 1] - c->h_sliceOff[s];
        for (int j = 0; j < w; j++)
            for (int l = 0; l < 32; l++) {
                int p = s * 32 + l;
                if (j >= c->h_rowNAll[p]) continue;
                size_t sl = ((size_t)c->h_sliceOff[s] + j) * 32 + l;
                c->h_gfid[sl] = ref2gf[c->h_meta[sl] >> 2];
            }
    }
    // face geometry in GPU face order
    const int NFG = c->NFG;
    std::vector<double> geo((size_t)NG * NFG);
    for (int g = 0; g < NFG; g++) {
        int f = c->h_gf2ref[g];
        geo[(size_t)G_SFX * NFG + g] = Sf[3 * (size_t)f];
        geo[(size_t)G_SFY * NFG + g] = Sf[3 * (size_t)f + 1];
        geo[(size_t)G_SFZ * NFG + g] = Sf[3 * (size_t)f + 2];
        geo[(size_t)G_MAGSF * NFG + g] = magSf[f];
        geo[(size_t)G_W * NFG + g] = weights[f];
        geo[(size_t)G_NONORTH * NFG + g] = nonOrthDeltaCoeffs[f];
        geo[(size_t)G_DELTA * NFG + g] = deltaCoeffs[f];
    }
    // cell centres / volumes in position order
    const int NPH = c->NPH;
    std::vector<double> Cp3((size_t)3 * NPH, 0.0), Vp(NP, 1.0);
    for (int p = 0; p < NP; p++) {
        int i = c->pos2cell[p];
        if (i < 0) continue;
        for (int d = 0; d < 3; d++) Cp3[(size_t)d * NPH + p] = C[3 * (size_t)i + d];
        Vp[p] = V[i];
    }
    // boundary faces
    std::vector<int> bfOwnerPos(NB, -1);
    std::vector<double> bfGeo((size_t)4 * std::max(NB, 1), 0.0), dCoupled((size_t)3 * std::max(NB, 1), 0.0);
    for (int b = 0; b < NB; b++) {
        if (c->bfacePatch[b] < 0) continue;
        int f = F + b;
        bfOwnerPos[b] = c->cell2pos[owner[f]];
        for (int d = 0; d < 3; d++) bfGeo[(size_t)d * NB + b] = Sf[3 * (size_t)f + d];
        bfGeo[(size_t)3 * NB + b] = magSf[f];
    }
    // coupled deltas (cyclic: own delta - neighbour-patch delta; processor: exchanged below)
    std::vector<double> ownDelta((size_t)3 * std::max(NB, 1), 0.0);
    for (int pi = 0; pi < n_patches; pi++) {
        const icsb200_patch& pa = patches[pi];
        if (pa.kind != ICSB200_CYCLIC && pa.kind != ICSB200_PROCESSOR && pa.kind != ICSB200_CYCLICAMI) continue;
        for (int f = pa.start; f < pa.start + pa.size; f++)
            for (int d = 0; d < 3; d++) ownDelta[3 * (size_t)(f - F) + d] = Cf[3 * (size_t)f + d] - C[3 * (size_t)owner[f] + d];
    }
    std::vector<double> nbrDelta(ownDelta.size(), 0.0);
    for (int pi = 0; pi < n_patches; pi++) {
        const icsb200_patch& pa = patches[pi];
        if (pa.kind == ICSB200_CYCLIC) {
            const icsb200_patch& qa = patches[pa.nbr_patch];
            const bool rot = ics_is_rotational(pa);   // cyclicFvPatch::delta(): patchD - transform(forwardT, nbrPatchD)
            const double* T = pa.forwardT;
            for (int i = 0; i < pa.size; i++) {
                const double* s3 = &ownDelta[3 * (size_t)(qa.start + i - F)];
                double* d3 = &nbrDelta[3 * (size_t)(pa.start + i - F)];
                if (rot) {
                    d3[0] = T[0] * s3[0] + T[1] * s3[1] + T[2] * s3[2];
                    d3[1] = T[3] * s3[0] + T[4] * s3[1] + T[5] * s3[2];
                    d3[2] = T[6] * s3[0] + T[7] * s3[1] + T[8] * s3[2];
                } else
                    for (int d = 0; d < 3; d++) d3[d] = s3[d];
            }
        } else if (pa.kind == ICSB200_CYCLICAMI) {
            // cyclicAMIFvPatch::delta(): patchD - interpolate(nbrPatch.coupledFvPatch::delta())
            const icsb200_patch& qa = patches[pa.nbr_patch];
            const AmiTable* t = nullptr;
            for (auto& pe : c->pendingAmi) if (pe.first == pi) t = &pe.second;
            for (int i = 0; i < pa.size; i++)
                for (int d = 0; d < 3; d++) {
                    double acc = 0.0;
                    for (int k = t->start[i]; k < t->start[i + 1]; k++) acc += t->weight[k] * ownDelta[3 * (size_t)(qa.start + t->face[k] - F) + d];
                    nbrDelta[3 * (size_t)(pa.start + i - F) + d] = acc;
                }
            if (ics_is_rotational(pa))   // ... - transform(forwardT, interpolated neighbour delta)
                for (int i = 0; i < pa.size; i++) {
                    double* d3 = &nbrDelta[3 * (size_t)(pa.start + i - F)];
                    const double* T = pa.forwardT;
                    const double v0 = d3[0], v1 = d3[1], v2 = d3[2];
                    d3[0] = T[0] * v0 + T[1] * v1 + T[2] * v2;
                    d3[1] = T[3] * v0 + T[4] * v1 + T[5] * v2;
                    d3[2] = T[6] * v0 + T[7] * v1 + T[8] * v2;
                }
        }
    }
    bool anyProc = false;
    for (int pi = 0; pi < n_patches; pi++) anyProc |= patches[pi].kind == ICSB200_PROCESSOR;
    if (NH > 0 && anyProc) {
        // exchange processor-patch deltas through NCCL (device staging)
        double *dS = nullptr, *dR = nullptr;
        CUDA_TRY(c, cudaMalloc((void**)&dS, sizeof(double) * 3 * NH));
        CUDA_TRY(c, cudaMalloc((void**)&dR, sizeof(double) * 3 * NH));
        std::vector<double> sendv((size_t)3 * NH), recvv((size_t)3 * NH);
        for (int pi = 0; pi < n_patches; pi++) {
            if (patches[pi].kind != ICSB200_PROCESSOR) continue;
            for (int i = 0; i < patches[pi].size; i++)
                for (int d = 0; d < 3; d++) sendv[3 * (size_t)(patchHaloStart[pi] + i) + d] = ownDelta[3 * (size_t)(patches[pi].start + i - F) + d];
        }
        CUDA_TRY(c, cudaMemcpy(dS, sendv.data(), sizeof(double) * 3 * NH, cudaMemcpyHostToDevice));
        ncclGroupStart();
        for (int pi = 0; pi < n_patches; pi++) {
            if (patches[pi].kind != ICSB200_PROCESSOR) continue;
            ncclSend(dS + 3 * (size_t)patchHaloStart[pi], 3 * (size_t)patches[pi].size, ncclDouble, patches[pi].nbr_rank, (ncclComm_t)c->nccl, c->stream);
            ncclRecv(dR + 3 * (size_t)patchHaloStart[pi], 3 * (size_t)patches[pi].size, ncclDouble, patches[pi].nbr_rank, (ncclComm_t)c->nccl, c->stream);
        }
        ncclGroupEnd();
        CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        CUDA_TRY(c, cudaMemcpy(recvv.data(), dR, sizeof(double) * 3 * NH, cudaMemcpyDeviceToHost));
        cudaFree(dS); cudaFree(dR);
        for (int pi = 0; pi < n_patches; pi++) {
            if (patches[pi].kind != ICSB200_PROCESSOR) continue;
            for (int i = 0; i < patches[pi].size; i++)
                for (int d = 0; d < 3; d++) nbrDelta[3 * (size_t)(patches[pi].start + i - F) + d] = recvv[3 * (size_t)(patchHaloStart[pi] + i) + d];
        }
    }
    for (int b = 0; b < NB; b++)
        for (int d = 0; d < 3; d++) dCoupled[(size_t)d * NB + b] = ownDelta[3 * (size_t)b + d] - nbrDelta[3 * (size_t)b + d];

    // ---- uploads
    int r = 0;
    r |= devUpload(c, &c->d_pos2cell, c->pos2cell);
    r |= devUpload(c, &c->d_cell2pos, c->cell2pos);
    r |= devUpload(c, &c->d_sliceOff, c->h_sliceOff);
    r |= devUpload(c, &c->d_rowNLow, c->h_rowNLow);
    r |= devUpload(c, &c->d_rowNInt, c->h_rowNInt);
    r |= devUpload(c, &c->d_rowNAll, c->h_rowNAll);
    r |= devUpload(c, &c->d_col, c->h_col);
    r |= devUpload(c, &c->d_meta, c->h_meta);
    r |= devUpload(c, &c->d_gfid, c->h_gfid);
    r |= devUpload(c, &c->d_geo, geo);
    r |= devUpload(c, &c->d_dCoupled, dCoupled);
    r |= devUpload(c, &c->d_V, Vp);
    r |= devUpload(c, &c->d_C, Cp3);
    if (c->tileMode && c->tileTma) {
        for (int p2 = 0; p2 < NP && c->tileTma; p2++) {
            if (c->pos2cell[p2] < 0) continue;
            if (c->h_rowNLow[p2] > 3 || c->h_rowNInt[p2] - c->h_rowNLow[p2] > 3 || c->h_rowNInt[p2] > 8) c->tileTma = false;
        }
    }
    if (c->tileMode) {
        r |= devUpload(c, &c->d_tileStart, tileStart);
        r |= devUpload(c, &c->d_tileFPtr, tileFPtr);
        r |= devUpload(c, &c->d_tileFLev, tileFLev);
        r |= devUpload(c, &c->d_tileRPtr, tileRPtr);
        r |= devUpload(c, &c->d_tileRLev, tileRLev);
        r |= devUpload(c, &c->d_tileRRows, tileRRows);
        std::vector<int> sliceTile(NP / 32, 0);
        c->tileMaxRows = 32;
        for (int t = 0; t < c->nTiles; t++) {
            for (int s = tileStart[t] / 32; s < tileStart[t + 1] / 32; s++) sliceTile[s] = t;
            c->tileMaxRows = std::max(c->tileMaxRows, tileStart[t + 1] - tileStart[t]);
        }
        r |= devUpload(c, &c->d_sliceTile, sliceTile);
        if (c->tileTma) {
            // per-row intra-tile levels of both sweeps, so the tile kernel can stage them with one bulk copy
            std::vector<int> rowLevF(NP, -1), rowLevR(NP, -1), nLevF2(c->nTiles, 0), nLevR2(c->nTiles, 0);
            for (int t = 0; t < c->nTiles; t++) {
                const int t0 = tileStart[t];
                const int nf = tileFPtr[t + 1] - tileFPtr[t] - 1, nr = tileRPtr[t + 1] - tileRPtr[t] - 1;
                nLevF2[t] = nf; nLevR2[t] = nr;
                for (int L = 0; L < nf; L++)
                    for (int rr2 = tileFLev[tileFPtr[t] + L]; rr2 < tileFLev[tileFPtr[t] + L + 1]; rr2++) rowLevF[t0 + rr2] = L;
                for (int L = 0; L < nr; L++)
                    for (int i2 = tileRLev[tileRPtr[t] + L]; i2 < tileRLev[tileRPtr[t] + L + 1]; i2++) rowLevR[t0 + tileRRows[t0 + i2]] = L;
            }
            r |= devUpload(c, &c->d_rowLevF, rowLevF);
            r |= devUpload(c, &c->d_rowLevR, rowLevR);
            r |= devUpload(c, &c->d_tileNLevF, nLevF2);
            r |= devUpload(c, &c->d_tileNLevR, nLevR2);
            // per (tile, sweep): t0, nRows, nLev, -, then per slice (2): first entry offset, first staged entry, staged entries,
            // staged column entries — everything the producer needs to issue a tile's bulk copies with one 64-byte read
            std::vector<int> descF((size_t)16 * c->nTiles, 0), descR((size_t)16 * c->nTiles, 0);
            for (int t = 0; t < c->nTiles; t++)
                for (int sw = 0; sw < 2; sw++) {
                    int* d = (sw == 0 ? descF.data() : descR.data()) + (size_t)16 * t;
                    const int t0 = tileStart[t], nRows = tileStart[t + 1] - t0;
                    d[0] = t0; d[1] = nRows; d[2] = sw == 0 ? nLevF2[t] : nLevR2[t];
                    for (int sl = 0; sl < nRows / 32 && sl < 2; sl++) {
                        const int s2 = t0 / 32 + sl;
                        int fHi = 0, rLo = 1 << 20, rHi = 0;
                        for (int l = 0; l < 32; l++) {
                            const int p2 = s2 * 32 + l;
                            fHi = std::max(fHi, c->h_rowNLow[p2]);
                            if (c->h_rowNInt[p2] > c->h_rowNLow[p2]) { rLo = std::min(rLo, c->h_rowNLow[p2]); rHi = std::max(rHi, c->h_rowNInt[p2]); }
                        }
                        if (rHi == 0) rLo = 0;
                        const int lo = sw == 0 ? 0 : rLo, hi = sw == 0 ? fHi : rHi;
                        d[4 + sl] = c->h_sliceOff[s2];
                        d[6 + sl] = lo;
                        d[8 + sl] = std::max(0, std::min(hi - lo, 3));
                        d[10 + sl] = std::min(c->h_sliceOff[s2 + 1] - c->h_sliceOff[s2], 8);
                    }
                }
            r |= devUpload(c, &c->d_tileDescF, descF);
            r |= devUpload(c, &c->d_tileDescR, descR);
        }
    }
    c->blkMode = false;
    if (c->tileMode && blkWanted) {
        // ---- block tiles (k_lusgs_blk, lusgs_blk.cu): everything a tile's sweep needs besides the 5x5 blocks, laid out so the
        // kernel's producer warp can fetch it with a handful of 16-byte-aligned bulk copies:
        //  blkTab  per tile [16 descriptor ints][level starts][forward halo positions][reverse halo positions][forward flags to
        //          wait for][reverse flags][first entry of each slice][first reverse-staged entry of each slice], every section
        //          padded to 4 ints (BT_* in lusgs_blk.cu);  blkIdx per tile {table offset, table length, t0, rows}
        //  blkInfo per sweep and row, 64 bits: for neighbour t = 0..2 (forward: ascending lower entries, reverse: descending
        //          upper entries) 10 bits local index (row of the tile, or ICS_BLK_MR + halo slot), 2 bits staged block slot
        //          (3 = not staged: read from global), 3 bits entry index; bits 45..46 = number of neighbours
        const int nT = c->nTiles;
        bool ok = true;
        std::vector<int> sliceTile2(NP / 32, 0);
        for (int t = 0; t < nT; t++) for (int s = tileStart[t] / 32; s < tileStart[t + 1] / 32; s++) sliceTile2[s] = t;
        std::vector<int> sFwdHi(nSlices, 0), sRevLo(nSlices, 0), sRevHi(nSlices, 0);
        for (int s2 = 0; s2 < nSlices; s2++) {
            int fHi = 0, rLo = 1 << 20, rHi = 0;
            for (int l = 0; l < 32; l++) {
                const int p2 = s2 * 32 + l;
                fHi = std::max(fHi, c->h_rowNLow[p2]);
                if (c->h_rowNInt[p2] > c->h_rowNLow[p2]) { rLo = std::min(rLo, c->h_rowNLow[p2]); rHi = std::max(rHi, c->h_rowNInt[p2]); }
            }
            if (rHi == 0) rLo = 0;
            sFwdHi[s2] = fHi; sRevLo[s2] = rLo; sRevHi[s2] = rHi;
        }
        std::vector<int> tab, idx((size_t)4 * nT, 0);
        std::vector<unsigned long long> info((size_t)2 * NP, 0ull);
        auto slotOf = [&](int p, int j) { return ((size_t)c->h_sliceOff[p / 32] + j) * 32 + (p % 32); };
        auto pad4 = [&]() { while (tab.size() % 4) tab.push_back(0); };
        std::vector<int> tmp[2], tls[2];
        for (int t = 0; t < nT && ok; t++) {
            const int t0 = tileStart[t], t1 = tileStart[t + 1];
            const int nLev = tileFPtr[t + 1] - tileFPtr[t] - 1;
            const int nSl = (t1 - t0) / 32;
            if (t1 - t0 > ICS_BLK_MR || nLev > ICS_BLK_MAXLEV || nLev < 1) { ok = false; break; }
            // the slices one level touches must fit in the block ring together (lusgs_blk.cu)
            for (int L = 0; L < nLev; L++) if (tileFLev[tileFPtr[t] + L + 1] - tileFLev[tileFPtr[t] + L] > ICS_BLK_MAXLW) ok = false;
            int unstaged[2] = {0, 0};  // the sweep has rows whose block lies outside the slice's staged entry range
            for (int sw = 0; sw < 2 && ok; sw++) {
                tmp[sw].clear();
                for (int p = t0; p < t1; p++) {
                    if (c->pos2cell[p] < 0) continue;
                    const int nLow = c->h_rowNLow[p], nInt = c->h_rowNInt[p];
                    if (nLow > 3 || nInt - nLow > 3) { ok = false; break; }
                    for (int j = sw == 0 ? 0 : nLow; j < (sw == 0 ? nLow : nInt); j++) {
                        const int q = c->h_col[slotOf(p, j)];
                        if (q >= t0 && q < t1) continue;
                        if (sw == 0 ? q >= t0 : q < t1) ok = false;  // the tile order must be a topological order
                        tmp[sw].push_back(q);
                    }
                }
                std::sort(tmp[sw].begin(), tmp[sw].end());
                tmp[sw].erase(std::unique(tmp[sw].begin(), tmp[sw].end()), tmp[sw].end());
                if ((int)tmp[sw].size() > ICS_BLK_MH - 1) ok = false;  // the last slot of the tile's vector is the zero the absent neighbours point to
                tls[sw].clear();
                // column mode: the tile swept just before this one by the same CTA (previous chunk of the column in the forward
                // sweep, next chunk in the reverse sweep) hands its values over in shared memory — no flag to wait for
                const int pred = sw == 0 ? t - 1 : t + 1;
                const bool hasPred = !tileCol.empty() && pred >= 0 && pred < nT && tileCol[pred] == tileCol[t];
                for (int q : tmp[sw]) {
                    const int tq = sliceTile2[q >> 5];
                    if (hasPred && tq == pred) continue;
                    tls[sw].push_back(tq + (sw == 0 ? 0 : nT));  // flag index: forward t, reverse nT + t
                }
                std::sort(tls[sw].begin(), tls[sw].end());
                tls[sw].erase(std::unique(tls[sw].begin(), tls[sw].end()), tls[sw].end());
                if (sw == 1) tls[sw].push_back(t);  // the reverse sweep starts from the tile's own forward values
                if ((int)tls[sw].size() > ICS_BLK_MAXDEP) ok = false;
                for (int p = t0; p < t1 && ok; p++) {
                    if (c->pos2cell[p] < 0) continue;
                    const int nLow = c->h_rowNLow[p], nInt = c->h_rowNInt[p];
                    const int n = sw == 0 ? nLow : nInt - nLow;
                    const int s2 = p / 32;
                    const int lo = sw == 0 ? 0 : sRevLo[s2];
                    const int cnt = sw == 0 ? std::min(sFwdHi[s2], ICS_BLK_SE) : std::max(0, std::min(sRevHi[s2] - sRevLo[s2], ICS_BLK_SE));
                    unsigned long long w = (unsigned long long)n << 45;
                    // absent neighbours (k >= n): local index of the zero slot, block slot 3 (the kernel's zero block)
                    for (int k = n; k < 3; k++) w |= (unsigned long long)((ICS_BLK_MR + ICS_BLK_MH - 1) | (3 << 10)) << (15 * k);
                    for (int k = 0; k < n; k++) {
                        const int j = sw == 0 ? k : nInt - 1 - k;
                        const int q = c->h_col[slotOf(p, j)];
                        int li;
                        if (q >= t0 && q < t1) li = q - t0;
                        else li = ICS_BLK_MR + (int)(std::lower_bound(tmp[sw].begin(), tmp[sw].end(), q) - tmp[sw].begin());
                        const int js = j - lo;
                        const int code = (js >= 0 && js < cnt) ? js : 3;
                        if (code == 3) unstaged[sw] = 1;
                        if (j > 7) ok = false;
                        w |= (unsigned long long)(li | (code << 10) | (j << 12)) << (15 * k);
                    }
                    info[(size_t)sw * NP + p] = w;
                }
            }
            if (!ok) break;
            pad4();
            const int off = (int)tab.size();
            tab.resize(off + 16, 0);
            auto section = [&](const std::vector<int>& v) { const int o = (int)tab.size() - off; tab.insert(tab.end(), v.begin(), v.end()); pad4(); return o; };
            std::vector<int> lv(tileFLev.begin() + tileFPtr[t], tileFLev.begin() + tileFPtr[t + 1]);
            std::vector<int> so(c->h_sliceOff.begin() + t0 / 32, c->h_sliceOff.begin() + t0 / 32 + nSl), rl(sRevLo.begin() + t0 / 32, sRevLo.begin() + t0 / 32 + nSl);
            int d[16] = {0};
            // column mode: halo entries that lie in the tile swept just before this one become -(row of that tile) - 1
            int colBits = 0;
            for (int sw = 0; sw < 2; sw++) {
                const int pred = sw == 0 ? t - 1 : t + 1;
                if (tileCol.empty() || pred < 0 || pred >= nT || tileCol[pred] != tileCol[t]) continue;
                colBits |= 1 << sw;
                for (int& q : tmp[sw]) if (q >= tileStart[pred] && q < tileStart[pred + 1]) q = -(q - tileStart[pred]) - 1;
            }
            d[0] = t0; d[1] = t1 - t0; d[2] = colBits; d[3] = nLev;
            d[4] = section(lv);
            d[5] = section(tmp[0]); d[6] = (int)tmp[0].size();
            d[7] = section(tmp[1]); d[8] = (int)tmp[1].size();
            d[9] = section(tls[0]); d[10] = (int)tls[0].size();
            d[11] = section(tls[1]); d[12] = (int)tls[1].size();
            d[13] = section(so); d[14] = section(rl); d[15] = nSl | (unstaged[0] << 16) | (unstaged[1] << 17);
            for (int k = 0; k < 16; k++) tab[off + k] = d[k];
            idx[(size_t)4 * t] = off; idx[(size_t)4 * t + 1] = (int)tab.size() - off; idx[(size_t)4 * t + 2] = t0; idx[(size_t)4 * t + 3] = t1 - t0;
            if ((int)tab.size() - off > ICS_BLK_TAB) ok = false;
        }
        if (ok) {
            // bulk-copy ranges per sweep and slice: first staged entry (absolute entry index) and number of staged entries
            std::vector<int> stage((size_t)4 * nSlices, 0);
            for (int s2 = 0; s2 < nSlices; s2++) {
                stage[2 * (size_t)s2] = c->h_sliceOff[s2];
                stage[2 * (size_t)s2 + 1] = std::min(sFwdHi[s2], ICS_BLK_SE);
                stage[2 * ((size_t)nSlices + s2)] = c->h_sliceOff[s2] + sRevLo[s2];
                stage[2 * ((size_t)nSlices + s2) + 1] = std::max(0, std::min(sRevHi[s2] - sRevLo[s2], ICS_BLK_SE));
            }
            r |= devUpload(c, &c->d_blkTab, tab);
            r |= devUpload(c, &c->d_blkIdx, idx);
            r |= devUpload(c, &c->d_blkInfo, info);
            r |= devUpload(c, &c->d_blkStage, stage);
            {
                // columns = units of work a CTA draws: [first tile, ...) per column; without column mode every tile is its own column
                std::vector<int> colStart;
                for (int t = 0; t < nT; t++) if (tileCol.empty() || t == 0 || tileCol[t] != tileCol[t - 1]) colStart.push_back(t);
                c->nBlkCols = (int)colStart.size();
                colStart.push_back(nT);
                r |= devUpload(c, &c->d_blkCol, colStart);
            }
            r |= devAlloc(c, &c->d_blkFlag, (size_t)2 * nT + 2);   // + the two ticket counters (lusgs_blk.cu)
            if (!r) CUDA_TRY(c, cudaMemset(c->d_blkFlag, 0, sizeof(int) * (2 * nT + 2)));
            c->blkEpoch = 0;
            c->blkMode = true;
        }
    }
    r |= devUpload(c, &c->d_bfOwnerPos, bfOwnerPos);
    r |= devUpload(c, &c->d_bfPatch, c->bfacePatch);
    {
        std::vector<int> bfKind(NB, -1);
        for (int b = 0; b < NB; b++) if (c->bfacePatch[b] >= 0) bfKind[b] = patches[c->bfacePatch[b]].kind;
        r |= devUpload(c, &c->d_bfKind, bfKind);
    }
    r |= devUpload(c, &c->d_bfGeo, bfGeo);
    if (r) return r;
    // the big host index arrays are no longer needed except the row bookkeeping used by matrix get/set
    // processor patches: send lists
    for (int pi = 0; pi < n_patches; pi++) {
        if (patches[pi].kind != ICSB200_PROCESSOR) continue;
        ProcPatchDev pp;
        pp.nbrRank = patches[pi].nbr_rank; pp.size = patches[pi].size; pp.haloStart = patchHaloStart[pi]; pp.d_sendPos = nullptr;
        std::vector<int> sp(pp.size);
        for (int i = 0; i < pp.size; i++) sp[i] = c->cell2pos[owner[patches[pi].start + i]];
        r |= devUpload(c, &pp.d_sendPos, sp);
        c->procs.push_back(pp);
    }
    for (auto& ro : c->rots) { cudaFree(ro.d_srcPos); if (ro.d_lagSrc) cudaFree(ro.d_lagSrc); }
    c->rots.clear();
    {
        // viscous terms across a rotational pair need the neighbour CELL (tauMC is rotated as a cell tensor) and forwardT
        std::vector<int> nbrPos((size_t)std::max(NB, 1), -1);
        std::vector<double> prot((size_t)10 * std::max(n_patches, 1), 0.0);
        for (int pi = 0; pi < n_patches; pi++) {
            if (!ics_is_rotational(patches[pi])) continue;
            prot[(size_t)10 * pi] = 1.0;
            for (int k = 0; k < 9; k++) prot[(size_t)10 * pi + 1 + k] = patches[pi].forwardT[k];
            if (patches[pi].kind == ICSB200_CYCLIC)
                for (int i = 0; i < patches[pi].size; i++) nbrPos[patches[pi].start + i - F] = c->cell2pos[owner[patches[patches[pi].nbr_patch].start + i]];
        }
        r |= devUpload(c, &c->d_bfNbrPos, nbrPos);
        r |= devUpload(c, &c->d_patchRot, prot);
    }
    for (int pi = 0; pi < n_patches; pi++) {
        if (!localHalo(pi)) continue;
        RotPatchDev ro{};
        ro.size = patches[pi].size; ro.haloStart = patchHaloStart[pi]; ro.d_srcPos = nullptr; ro.d_lagSrc = nullptr; ro.nLag = 0;
        ro.rotate = ics_is_rotational(patches[pi]);
        std::memcpy(ro.T, patches[pi].forwardT, sizeof(ro.T));
        std::vector<int> sp(ro.size);
        for (int i = 0; i < ro.size; i++) sp[i] = c->cell2pos[owner[patches[patches[pi].nbr_patch].start + i]];
        r |= devUpload(c, &ro.d_srcPos, sp);
        if (const std::vector<double>* lw = lagOf(pi)) {
            // instance J's copy of the neighbour cell: the mesh is n_instants copies of NC cells, instance-major (icsb200_hb_set)
            ro.nLag = (int)lw->size();
            const int NC = N / ro.nLag;
            for (int J = 0; J < ro.nLag; J++) ro.lagW.w[J] = (*lw)[J];
            std::vector<int> ls((size_t)ro.nLag * ro.size);
            for (int J = 0; J < ro.nLag; J++)
                for (int i = 0; i < ro.size; i++) ls[(size_t)J * ro.size + i] = c->cell2pos[J * NC + owner[patches[patches[pi].nbr_patch].start + i] % NC];
            r |= devUpload(c, &ro.d_lagSrc, ls);
        }
        c->rots.push_back(ro);
    }
    for (auto& am : c->amis) { cudaFree(am.d_start); cudaFree(am.d_srcPos); cudaFree(am.d_w); }
    c->amis.clear();
    for (int pi = 0; pi < n_patches; pi++) {
        if (patches[pi].kind != ICSB200_CYCLICAMI) continue;
        const AmiTable* t = nullptr;
        for (auto& pe : c->pendingAmi) if (pe.first == pi) t = &pe.second;
        AmiPatchDev am{};
        am.size = patches[pi].size; am.haloStart = patchHaloStart[pi];
        am.rot = ics_is_rotational(patches[pi]);
        std::memcpy(am.T, patches[pi].forwardT, sizeof(am.T));
        std::vector<int> sp(t->face.size());
        for (size_t k = 0; k < sp.size(); k++) sp[k] = c->cell2pos[owner[patches[patches[pi].nbr_patch].start + t->face[k]]];
        r |= devUpload(c, &am.d_start, t->start);
        r |= devUpload(c, &am.d_srcPos, sp);
        r |= devUpload(c, &am.d_w, t->weight);
        c->amis.push_back(am);
    }
    {
        // per-boundary-face view of the AMI stencils: tauMC's patchNeighbourField interpolates the neighbour CELLS' tensors (k_visc)
        std::vector<int> bfStart((size_t)NB + 1, 0), allSrc;
        std::vector<double> allW;
        for (int b = 0; b < NB; b++) {
            const int pi = c->bfacePatch[b];
            bfStart[b] = (int)allSrc.size();
            if (pi < 0 || patches[pi].kind != ICSB200_CYCLICAMI) continue;
            const AmiTable* t = nullptr;
            for (auto& pe : c->pendingAmi) if (pe.first == pi) t = &pe.second;
            const int i = F + b - patches[pi].start;
            for (int k = t->start[i]; k < t->start[i + 1]; k++) {
                allSrc.push_back(c->cell2pos[owner[patches[patches[pi].nbr_patch].start + t->face[k]]]);
                allW.push_back(t->weight[k]);
            }
        }
        bfStart[NB] = (int)allSrc.size();
        if (allSrc.empty()) { allSrc.push_back(0); allW.push_back(0.0); }
        r |= devUpload(c, &c->d_bfAmiStart, bfStart);
        r |= devUpload(c, &c->d_amiAllSrc, allSrc);
        r |= devUpload(c, &c->d_amiAllW, allW);
    }
    if (NH > 0) {
        r |= devAlloc(c, &c->d_sendBuf, (size_t)NH * 40);
        r |= devAlloc(c, &c->d_recvBuf, (size_t)NH * 40);
    }
    // halo centres are never read (coupled faces use dCoupled)
    // ---- BCs default: zeroGradient on physical patches
    for (auto& hb : c->h_bc) for (int fl = 0; fl < 3; fl++) if (hb.prmFace[fl]) cudaFree((void*)hb.prmFace[fl]);
    c->h_bc.assign(n_patches, BCDev{});
    for (int pi = 0; pi < n_patches; pi++) c->h_bc[pi].bstart = patches[pi].start - F;
    for (int pi = 0; pi < n_patches; pi++) {
        int k = (patches[pi].kind == ICSB200_CYCLIC || patches[pi].kind == ICSB200_PROCESSOR || patches[pi].kind == ICSB200_CYCLICAMI) ? ICSB200_BC_COUPLED
                : patches[pi].kind == ICSB200_EMPTY ? ICSB200_BC_EMPTY : ICSB200_BC_ZEROGRADIENT;
        for (int fl = 0; fl < 3; fl++) c->h_bc[pi].kind[fl] = k;
    }
    r |= devUpload(c, &c->d_bc, c->h_bc);
    // ---- field storage
    const size_t NX = c->NX;
    r |= devAlloc(c, &c->d_fields, (size_t)Q_COUNT * NX);
    r |= devAlloc(c, &c->d_grad, (size_t)NQ * 3 * NPH);
    r |= devAlloc(c, &c->d_rdt, (size_t)NP);
    r |= devAlloc(c, &c->d_co, (size_t)NP);
    r |= devAlloc(c, &c->d_ddtCoeff, (size_t)NP);
    r |= devAlloc(c, &c->d_Wold, (size_t)5 * NP);
    r |= devAlloc(c, &c->d_Wold2, (size_t)5 * NP);
    r |= devAlloc(c, &c->d_Wprev, (size_t)5 * NPH);
    r |= devAlloc(c, &c->d_src, (size_t)5 * NPH);
    r |= devAlloc(c, &c->d_dW, (size_t)5 * NPH);
    r |= devAlloc(c, &c->d_bad, (size_t)NPH);
    r |= devAlloc(c, &c->d_phiB, (size_t)std::max(NB, 1));
    r |= devAlloc(c, &c->d_vic, (size_t)9 * std::max(NB, 1));
    r |= devAlloc(c, &c->d_offd, nE32 * 25);
    r |= devAlloc(c, &c->d_diag, (size_t)25 * NP);
    r |= devAlloc(c, &c->d_rD, (size_t)NP);
    r |= devAlloc(c, &c->d_w, (size_t)5 * NPH);
    r |= devAlloc(c, &c->d_x, (size_t)5 * NPH);
    if (r) return r;
    CUDA_TRY(c, cudaMemset(c->d_fields, 0, sizeof(double) * Q_COUNT * NX));
    CUDA_TRY(c, cudaMemset(c->d_grad, 0, sizeof(double) * NQ * 3 * NPH));
    CUDA_TRY(c, cudaMemset(c->d_offd, 0, sizeof(double) * nE32 * 25));
    CUDA_TRY(c, cudaMemset(c->d_diag, 0, sizeof(double) * 25 * NP));
    CUDA_TRY(c, cudaMemset(c->d_src, 0, sizeof(double) * 5 * NPH));
    CUDA_TRY(c, cudaMemset(c->d_dW, 0, sizeof(double) * 5 * NPH));
    CUDA_TRY(c, cudaMemset(c->d_Wprev, 0, sizeof(double) * 5 * NPH));
    CUDA_TRY(c, cudaMemset(c->d_w, 0, sizeof(double) * 5 * NPH));
    CUDA_TRY(c, cudaMemset(c->d_x, 0, sizeof(double) * 5 * NPH));
    CUDA_TRY(c, cudaMemset(c->d_rdt, 0, sizeof(double) * NP));
    CUDA_TRY(c, cudaMemset(c->d_rD, 0, sizeof(double) * NP));
    CUDA_TRY(c, cudaMemset(c->d_co, 0, sizeof(double) * NP));
    CUDA_TRY(c, cudaMemset(c->d_ddtCoeff, 0, sizeof(double) * NP));
    CUDA_TRY(c, cudaMemset(c->d_phiB, 0, sizeof(double) * std::max(NB, 1)));
    CUDA_TRY(c, cudaMemset(c->d_vic, 0, sizeof(double) * 9 * std::max(NB, 1)));
    CUDA_TRY(c, cudaMemset(c->d_bad, 0, sizeof(int) * NPH));
    c->lusgsGrid = 0;
    c->lusgsTileGrid = 0;
    c->meshSet = true;
    c->stateSet = c->matrixSet = c->fluxValid = false;
    c->reconValid = false;
    c->hbNO = 1;  // a new mesh drops the Harmonic Balance setup (icsb200_hb_set must follow mesh_set)
    c->pendingLag.clear();               // consumed: a later mesh needs its own icsb200_phaselag_set calls
    r |= devAlloc(c, &c->d_mrfFace, 0);  // ... and the MRF fields (icsb200_mrf_set must follow mesh_set)
    r |= devAlloc(c, &c->d_mrfOmega, 0);
    r |= devAlloc(c, &c->d_transport, 0);
    c->srcMrfApplied = false;
    return r;
}
