// common.cuh — context, device layout and launch helpers of libicsb200 (sm_100a).
//
// Device layout (private; see DESIGN.md "Data layout in HBM"):
//  * cells are renumbered into POSITIONS in LU-SGS sweep order (setup.cu): on hex-like meshes by column, tile (chunk of
//    hyperplanes of a column) and forward level inside the tile (lusgs_blk.cu), otherwise by forward level (longest path in
//    the owner<neighbour DAG, lusgs.C:130-159), ties by cell id; tiles / levels are padded to a multiple of 32 so a warp never
//    straddles one.  All cell vectors are SoA of length NPH = NP (+ halo slots + boundary-face slots).
//  * the coupled matrix (coupledMatrix.H: 9 LDU sub-blocks) is stored row-wise as sliced-ELL with one 5x5
//    block per face of the row (SELL-32, entry-major SoA): value (e, k, lane) at ((sliceOff[s]+j)*25+k)*32+lane.
//    Row entries are the faces of the cell in ascending reference face id = [faces where the cell is the
//    neighbour][faces the cell owns][boundary faces], which is exactly the order in which the reference's
//    face loops touch that cell (negSumDiag, Amul, fvc::div, gaussGrad) — so per-row accumulation in entry
//    order reproduces the reference's rounding without atomics or colouring.
//  * block variable order: (rho, rhoUx, rhoUy, rhoUz, rhoE).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "icsb200.h"

#define ICS_SMALL 1e-15
#define ICS_VSMALL 1e-300
#define ICS_ROOTVSMALL 1e-150
#define ICS_GREAT 1e15
#define ICS_VGREAT 1e300
#define ICS_TILE_MAXROWS 1024  // rows a LU-SGS tile may hold (shared-memory staging of 5 doubles per row)
// block tiles (k_lusgs_blk): rows per tile, out-of-tile neighbours per sweep, intra-tile levels, tiles waited for, staged entries per slice
#define ICS_BLK_MR 256
#define ICS_BLK_MH 160
#define ICS_BLK_MAXLEV 63
#define ICS_BLK_MAXDEP 32
#define ICS_BLK_SE 3
#define ICS_BLK_MAXLW 64   // widest intra-tile level (three consecutive levels must fit in the block ring together)
#define ICS_BLK_TAB 512    // ints of a tile's table

constexpr int NQ = 8;   // reconstructed scalars: rho, p, Ux, Uy, Uz, cR, E, H
constexpr int NG = 7;   // geometry doubles per face (SoA over GPU face ids)
enum { G_SFX = 0, G_SFY, G_SFZ, G_MAGSF, G_W, G_NONORTH, G_DELTA };
// entry types (low 2 bits of meta); face id = meta >> 2
enum { ET_LOWER = 0, ET_UPPER = 1, ET_COUPLED = 2, ET_PHYS = 3 };
// cell field ids (SoA arrays of length NPH)
enum {
    Q_RHO = 0, Q_P, Q_UX, Q_UY, Q_UZ, Q_CR, Q_E, Q_H,  // the NQ reconstructed scalars (order matters)
    Q_C,                                                 // sqrt(gamma/psi) (lambda)
    Q_EC,                                                // eCalc = rhoE/rho - 0.5 |U|^2 (viscous residual, residualsUpdate.H:40)
    Q_T, Q_PSI,
    Q_W0, Q_W1, Q_W2, Q_W3, Q_W4,                        // conserved rho, rhoU(3), rhoE
    Q_COUNT
};

struct BCDev {
    int kind[3];       // p, U, T
    double prm[3][8];
    // non-uniform entries (icsb200_bc_set_nonuniform): 8 doubles per face of the patch, or null; bstart = first boundary-face index
    const double* prmFace[3];
    int bstart;
    __host__ __device__ const double* P(int field, int b) const { return prmFace[field] ? prmFace[field] + (size_t)8 * (b - bstart) : prm[field]; }
};

// timer classes
enum {
    TM_BC = 0, TM_PRIM, TM_GRAD, TM_FLUX, TM_JAC, TM_SPMV, TM_LUSGS, TM_JACOBI, TM_VEC, TM_RED, TM_UPDATE, TM_PERM, TM_HALO, TM_COUNT
};
static const char* const kTimerNames[TM_COUNT] = {"bc", "primitives", "gradient", "flux_residual", "jacobian", "spmv", "lusgs",
                                                  "jacobi", "vector_ops", "reductions", "update", "permute", "halo"};

struct ProcPatchDev {
    int nbrRank, size;
    int haloStart;   // first halo slot (relative to NP)
    int* d_sendPos;  // positions of faceCells
};

// cyclicAMI patch on the device: halo slots [NP+haloStart, +size) receive sum_k w_k phi[srcPos_k] (CSR over the faces)
struct AmiPatchDev {
    int size, haloStart;
    bool rot;        // rotational pair: transform(forwardT, interpolated value) (cyclicAMIFvPatchField.C:171-203)
    double T[9];
    int* d_start;    // [size+1]
    int* d_srcPos;   // [nnz] positions of the neighbour patch's face cells
    double* d_w;     // [nnz]
};
// rotational cyclic patch on the device: halo slots [NP+haloStart, +size) receive transform(forwardT, phi[srcPos_i]) —
// vector triples rotated, scalars copied (cyclicFvPatchField::patchNeighbourField with doTransform())
struct LagWeights { double w[16]; };
struct RotPatchDev {
    int size, haloStart;
    int* d_srcPos;   // [size] positions of the neighbour patch's face cells
    double T[9];     // forwardT, row-major
    bool rotate;     // forwardT != I
    // phase lag (icsb200_phaselag_set): the listed fields take sum_J lagW[J] * value at d_lagSrc[J*size + i] (instance J's copy
    // of the neighbour cell); nLag = 0: plain cyclic
    int nLag;
    int* d_lagSrc;
    LagWeights lagW;
};
struct AmiTable { std::vector<int> start, face; std::vector<double> weight; };

struct icsb200_ctx {
    int device = 0, rank = 0, nRanks = 1;
    cudaStream_t stream = nullptr, commStream = nullptr;
    void* nccl = nullptr;  // ncclComm_t
    std::string err;
    long long launches = 0;
    int numSMs = 148;

    // ---- host mesh copy (reference numbering) ----
    int N = 0, F = 0, FT = 0, NB = 0;
    std::vector<int> owner, neighbour;
    std::vector<icsb200_patch> patches;
    std::vector<int> bfacePatch;  // [NB] patch of a boundary face (-1: empty)
    std::vector<int> bfNbrSlot;   // [NB] coupled faces: extended slot holding the patchNeighbourField (neighbour cell or halo slot); else -1
    int solutionD[3] = {1, 1, 1};
    bool meshSet = false, stateSet = false, matrixSet = false, fluxValid = false, thermoSet = false;

    // ---- positions ----
    int NP = 0, NH = 0, NPH = 0, NX = 0;  // positions, halo slots, NP+NH, NP+NH+NB (extended slots)
    int nSlices = 0;
    long long nEntries = 0;  // total slice entries (each = 32 lanes)
    std::vector<int> pos2cell, cell2pos;
    std::vector<int> h_sliceOff;
    std::vector<int> h_rowNLow, h_rowNInt, h_rowNAll;
    std::vector<int> h_col, h_meta, h_gfid;  // [nEntries*32]: neighbour slot, type | refFaceId<<2, GPU face id
    std::vector<int> h_gf2ref;               // [NFG] GPU face id -> reference face id
    int NFG = 0;                             // number of GPU faces (all non-empty faces)
    int *d_pos2cell = nullptr, *d_cell2pos = nullptr, *d_sliceOff = nullptr, *d_rowNLow = nullptr, *d_rowNInt = nullptr,
        *d_rowNAll = nullptr, *d_col = nullptr, *d_meta = nullptr, *d_gfid = nullptr;
    double* d_geo = nullptr;  // [NG*NFG] face geometry, SoA over GPU face ids (owner-side entry order)
    double* d_dCoupled = nullptr;  // [3*NB] delta vector of coupled boundary faces (SoA)
    double* d_V = nullptr;    // [NP] cell volumes (1 for padding)
    double* d_C = nullptr;    // [3*NPH] cell centres (SoA)
    // levels
    int nLevF = 0, nLevR = 0, maxWidth = 0;
    // blocked-wavefront schedule (tile mode)
    bool tileMode = false;
    bool tileTma = false;  // tiles of <= 64 rows swept by k_lusgs_tile_tma (thread = row, TMA-staged blocks, double-buffered)
    int nTiles = 0, nTileLevels = 0;
    int tileMaxRows = 0, lusgsTileGrid = 0;
    int* d_sliceTile = nullptr;  // [nSlices] tile of a slice
    // block-tile sweep (k_lusgs_blk, the default schedule when the mesh allows it; solver.cu "block tiles")
    bool blkMode = false;
    int* d_blkTab = nullptr;     // per-tile tables (setup.cu "block tiles")
    int* d_blkIdx = nullptr;     // [nTiles][4] table offset, table length, first position, rows
    unsigned long long* d_blkInfo = nullptr;  // [2][NP] packed per-row neighbour info of the forward / reverse sweep
    long long* d_blkProf = nullptr;           // optional per-CTA phase cycle counters (ICSB200_LUSGS_PROF)
    long long* d_blkTrace = nullptr;          // optional per-tile global-timer stamps (ICSB200_LUSGS_TRACE)
    int* d_blkStage = nullptr;   // [2][nSlices][2] per sweep and slice: first staged block entry, number of staged entries
    int* d_blkFlag = nullptr;    // [2*nTiles] completion epochs: forward sweep of tile t, reverse sweep of tile t
    int* d_blkCol = nullptr;     // [nBlkCols + 1] first tile of each column (the unit of work a CTA draws; lusgs_blk.cu)
    int nBlkCols = 0;
    int blkEpoch = 0;
    int *d_rowLevF = nullptr, *d_rowLevR = nullptr;      // [NP] intra-tile level of a row in the forward / reverse sweep (-1 padding)
    int *d_tileNLevF = nullptr, *d_tileNLevR = nullptr;  // [nTiles] number of intra-tile levels
    int *d_tileDescF = nullptr, *d_tileDescR = nullptr;  // [nTiles][16] bulk-copy descriptors of a tile's sweep (solver.cu TT_DESC_*)
    int *d_tileStart = nullptr, *d_tileFPtr = nullptr, *d_tileFLev = nullptr, *d_tileRPtr = nullptr, *d_tileRLev = nullptr, *d_tileRRows = nullptr;
    // boundary faces
    int* d_bfOwnerPos = nullptr;  // [NB] position of faceCell (-1 for empty)
    int* d_bfPatch = nullptr;     // [NB]
    int* d_bfKind = nullptr;      // [NB] fvPatch kind of the face's patch (-1: empty)
    double* d_bfGeo = nullptr;    // [4*NB] Sf(3), magSf  (SoA)
    BCDev* d_bc = nullptr;        // [nPatches]
    std::vector<BCDev> h_bc;
    double *d_phiB = nullptr;               // [NB] mass flux through boundary faces (lagged, for inletOutlet-type BCs)
    double *d_vic = nullptr;                // [5*NB] pVIC, uVIC(3), tVIC frozen at BC evaluation
    // processor patches
    std::vector<ProcPatchDev> procs;
    std::vector<AmiPatchDev> amis;                 // cyclicAMI patches (local weighted gathers into halo slots)
    std::vector<RotPatchDev> rots;                 // rotational cyclic patches (local gathers + rotation of the vector triples)
    int* d_bfNbrPos = nullptr;                     // [NB] rotational cyclic faces: position of the neighbour patch's face cell; else -1
    int* d_bfAmiStart = nullptr;                   // [NB+1] CSR offsets of the AMI stencil of a boundary face (empty range: not an AMI face)
    int* d_amiAllSrc = nullptr;                    // concatenated stencil positions / weights of all cyclicAMI patches (viscous terms)
    double* d_amiAllW = nullptr;
    double* d_patchRot = nullptr;                  // [10*nPatches] (rotational ? 1 : 0, forwardT[9]) per patch (viscous terms)
    std::vector<std::pair<int, AmiTable>> pendingAmi;  // icsb200_ami_set tables waiting for mesh_set
    std::vector<std::pair<int, std::vector<double>>> pendingLag;  // icsb200_phaselag_set rows waiting for mesh_set
    double *d_sendBuf = nullptr, *d_recvBuf = nullptr;

    // ---- thermo / schemes ----
    double R = 287, Cp = 1005, Cv = 718, gamma = 1.4, mu = 0, Pr = 1;
    icsb200_schemes sch{};
    double pseudoCoNum = 1;
    bool haveInitRes = false, havePrevRes = false, firstIter = true;
    icsb200_residuals initRes{}, prevRes{};
    int timeIndex = 0;

    // ---- fields: d_q[Q_COUNT] each [NX]; gradients d_grad [NQ*3] each [NPH] ----
    double* d_fields = nullptr;  // Q_COUNT * NX
    double* d_grad = nullptr;    // NQ*3 * NPH
    double *d_rdt = nullptr, *d_co = nullptr, *d_ddtCoeff = nullptr;  // [NP]
    double *d_Wold = nullptr, *d_Wold2 = nullptr, *d_Wprev = nullptr; // [5*NP] each
    double *d_src = nullptr, *d_dW = nullptr;                         // [5*NPH]
    double* d_faceFlux = nullptr;                                     // [5*NFG] face fluxes, GPU face order
    double* d_faceRecon = nullptr;                                    // [8*NFG] limited L/R states U_l U_r E_l E_r of every face (for k_jac)
    bool reconValid = false;                                          // d_faceRecon belongs to the current state and schemes
    double* d_gradE = nullptr;                                        // [3*NPH] gradient of eCalc (viscous runs only)
    double* d_visc = nullptr;                                         // [8*NP] per-row viscous divergences: lapU(3) divTau(3) divSigmaU lapE
    // MRF (icsb200_mrf_set): flux.MRFFaceVelocity() in GPU face order and flux.MRFOmega() by position; null = zero field
    double* d_mrfFace = nullptr;                                      // [NFG]
    double* d_mrfOmega = nullptr;                                     // [3*NP]
    bool srcMrfApplied = false;                                       // addMRFSource's Coriolis term is already in d_src
    // turbulence->muEff() / alphaEff() from the caller (icsb200_transport_set): [2][NX] over positions, halo and boundary
    // slots; null = the laminar constants mu and gamma mu / Pr
    double* d_transport = nullptr;
    int* d_bad = nullptr;                                              // [NPH] boundLocalTimeStep flags
    // ---- matrix ----
    double* d_offd = nullptr;  // [nEntries*25*32]
    double* d_diag = nullptr;  // [25*NP]
    double* d_rD = nullptr;    // [NP] lusgs rDiagCoeff
    double* d_invD = nullptr;  // [25*NP] Jacobi inverse blocks (lazily)
    bool rDValid = false, invDValid = false;
    // ---- solver work ----
    int mAlloc = 0;
    double* d_kry = nullptr;  // [m][5*NPH]
    double *d_w = nullptr, *d_x = nullptr;  // [5*NPH]
    double* d_scal = nullptr;  // device scalars
    double* h_scal = nullptr;  // pinned mirror
    double* d_partial = nullptr;
    unsigned int* d_counter = nullptr;
    unsigned int* d_barrier = nullptr;
    int lusgsGrid = 0;
    double* d_lusgsYZ = nullptr;  // [2][5*NPH] forward / reverse sweep values (sentinel protocol)
    int* d_lusgsHint = nullptr;   // [2*nSlices] publication hints
    int* d_sliceRange = nullptr;  // [3*nSlices] per slice: forward hi, reverse lo, reverse hi entry index (TMA staging ranges)
    bool lusgsTmaReady = false;
    long long* d_lusgsTrace = nullptr;
    int lusgsEpoch = 0;
    // ---- harmonic balance (hb.cu): nO time instances as one mesh of nO disconnected copies ----
    int hbNO = 1, hbNZones = 0;
    double* d_hbD = nullptr;      // [nZones][nO][nO]
    int* d_hbPeer = nullptr;      // [nO][NP] position of the same cell in instance K (-1 for padding rows)
    int* d_hbInst = nullptr;      // [NP] instance of a position (-1 padding)
    int* d_hbZone = nullptr;      // [NP] HB zone of the position's cell (-1 none)
    void* d_hbZonePrm = nullptr;  // [nZones] HBZonePrm
    double* d_hbInv = nullptr;    // [(5 nO)^2][NP/nO...] dense Jacobi inverses (lazily)
    double* d_hbWork = nullptr;
    std::vector<double> hbSInit, hbVInit, hbSFinal, hbVFinal, hbSInitPrev, hbVInitPrev;
    // staging
    double* d_stage = nullptr;
    size_t stageBytes = 0;
    double* h_pinned = nullptr;
    size_t pinnedBytes = 0;

    // timers
    bool timing = false;
    double tms[TM_COUNT] = {0};
    long long tcalls[TM_COUNT] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evA = nullptr, evB = nullptr;

    double* q(int id) const { return d_fields + (size_t)id * NX; }
    double* grad(int qi, int d) const { return d_grad + (size_t)(qi * 3 + d) * NPH; }
};

#define CUDA_TRY(ctx, call)                                                                            \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return ICSB200_ECUDA;                                                                      \
        }                                                                                              \
    } while (0)

inline int ics_fail(icsb200_ctx* c, int code, const std::string& msg) { c->err = msg; return code; }

// RAII-ish timer/launch accounting around one kernel launch
struct LaunchScope {
    icsb200_ctx* c;
    int cls;
    LaunchScope(icsb200_ctx* c_, int cls_) : c(c_), cls(cls_)
    {
        c->launches++;
        if (c->timing) cudaEventRecord(c->ev0, c->stream);
    }
    ~LaunchScope()
    {
        if (c->timing) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            c->tms[cls] += ms;
            c->tcalls[cls]++;
        }
    }
};

template <class T>
inline int devAlloc(icsb200_ctx* c, T** p, size_t n)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) return 0;
    CUDA_TRY(c, cudaMalloc((void**)p, n * sizeof(T)));
    return 0;
}
template <class T>
inline int devUpload(icsb200_ctx* c, T** p, const std::vector<T>& v)
{
    int r = devAlloc(c, p, v.size());
    if (r) return r;
    if (!v.empty()) CUDA_TRY(c, cudaMemcpy(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

inline int gridFor(long long n, int block) { return (int)((n + block - 1) / block); }

// ---- internal entry points across translation units ----
int ics_ensure_stage(icsb200_ctx* c, size_t bytes);
// cell-order AoS host array (nc comps) -> position-order SoA device arrays dst[k] = base + k*stride
int ics_upload_cells(icsb200_ctx* c, const double* host, int nc, double* dst, size_t stride);
int ics_download_cells(icsb200_ctx* c, double* host, int nc, const double* src, size_t stride);
int ics_eval_bc(icsb200_ctx* c, bool init);
int ics_primitives(icsb200_ctx* c);  // derived fields from p,T,U,rho (cells + boundary slots)
// vecMask: bit a set = arrays a, a+1, a+2 are the components of a vector (rotated on rotational cyclic patches)
// lagMask: bit a set = array a is one of the fields phase-lag patches mix over the time instances (rho p U cR E H c)
int ics_halo_fields(icsb200_ctx* c, double* base, size_t stride, int nArrays, unsigned vecMask = 0u, unsigned lagMask = 0u);
inline bool ics_is_rotational(const icsb200_patch& p)
{
    if (p.kind != ICSB200_CYCLIC && p.kind != ICSB200_CYCLICAMI) return false;
    for (int k = 0; k < 9; k++) if (std::fabs(p.forwardT[k] - (k % 4 == 0 ? 1.0 : 0.0)) > 1e-12) return true;
    return false;
}
int ics_gradients(icsb200_ctx* c);
int ics_flux_residual(icsb200_ctx* c, bool storeFaceFlux);
int ics_pseudo_ser(icsb200_ctx* c);
int ics_jacobian(icsb200_ctx* c, bool useStoredRdt);
int ics_rdiag(icsb200_ctx* c);
int ics_spmv(icsb200_ctx* c, const double* x, double* y, const double* b /*nullable: y = b - A x*/);
int ics_lusgs(icsb200_ctx* c, double* x);
int ics_lusgs_blk(icsb200_ctx* c, double* x);  // lusgs_blk.cu
int ics_jacobi_prepare(icsb200_ctx* c);
int ics_jacobi(icsb200_ctx* c, double* x);
int ics_gmres(icsb200_ctx* c, const icsb200_solver_controls* ctl, icsb200_residuals* res);
int ics_bound_local_dt(icsb200_ctx* c);
int ics_update(icsb200_ctx* c);
int ics_copy_prev(icsb200_ctx* c);
int ics_state_from_primitives(icsb200_ctx* c);
int ics_allreduce_max_int(icsb200_ctx* c, int* d, int n);
int ics_allreduce_max_double(icsb200_ctx* c, double* d, int n);
// hb.cu
int ics_mrf_source(icsb200_ctx* c);       // src(rhoU) -= (Omega ^ rho U) V, once per residual evaluation (jacobian.cu)
int ics_hb_source(icsb200_ctx* c);        // src += HB source (after the flux residual)
int ics_hb_diag(icsb200_ctx* c);          // diag += V D[J][J] (after the Jacobian)
int ics_hb_rdiag(icsb200_ctx* c);         // shared lusgs rDiagCoeff over all instances
int ics_hb_jacobi(icsb200_ctx* c, double* x);
struct HBSpmv { int nO; const int* peer; const int* inst; const int* zone; const double* D; const double* V; };
inline HBSpmv ics_hb_spmv_args(const icsb200_ctx* c) { return HBSpmv{c->hbNO, c->d_hbPeer, c->d_hbInst, c->d_hbZone, c->d_hbD, c->d_V}; }
