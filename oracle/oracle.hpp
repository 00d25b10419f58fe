// oracle.hpp — data structures of the CPU oracle.  TEST INFRASTRUCTURE ONLY.
//
// The oracle is a plain C++ restatement of ICSFoam's implicit pseudo-time iteration
// (SURVEY.md §8a/§8c).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load it; the product (icsfoam_b200/csrc) never does.
//
// PARITY UNPINNED: the reference ships no golden vectors and cannot be compiled here (OpenFOAM v2112
// is absent — SURVEY.md §0.9/§8c).  The arithmetic that lives in OpenFOAM itself (Gauss gradient, NVD/TVD
// limited interpolation, thermo, boundary conditions, LduMatrix) is restated from SURVEY.md Appendix A
// and pinned only by the analytic known-answer tests in tests/test_oracle_*.py; the formulas that live in /root/reference
// itself (flux schemes, Jacobian, pseudo time step, residual / update, viscous residual, GMRES + preconditioners) are
// cross-checked against independent numpy second readings (tests/test_*_second_reading.py).  How to pin it against a real
// ICSFoam build: tools/openfoam_golden/README.md.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../include/icsb200.h"

namespace orc {

using vecd = std::vector<double>;
using veci = std::vector<int>;

constexpr double SMALL = 1e-15, VSMALL = 1e-300, ROOTVSMALL = 1e-150, GREAT = 1e15, VGREAT = 1e300;

inline double sqr(double x) { return x * x; }
inline double pos0(double s) { return s >= 0 ? 1.0 : 0.0; }  // OpenFOAM pos()/pos0(): s >= 0
inline double neg(double s) { return s < 0 ? 1.0 : 0.0; }
inline double sign(double s) { return s >= 0 ? 1.0 : -1.0; }
inline double stabilise(double x, double y) { return x < 0 ? x - y : x + y; }

struct Patch {
    int kind, start, size, nbrRank, nbrPatch;
    double forwardT[9];
    bool rotational = false;  // cyclic with forwardT != I: patchNeighbourField = transform(forwardT, neighbour cell value)
    // cyclicAMI: CSR over the patch faces of (face index within the neighbour patch, weight)
    veci amiStart, amiFace;
    vecd amiWeight;
};

struct Mesh {
    int N = 0, F = 0, FT = 0, NB = 0;
    veci owner, neighbour;
    vecd Sf, magSf, w, deltaCoeffs, nonOrthDeltaCoeffs, C, V, Cf;
    std::vector<Patch> patches;
    int solutionD[3] = {1, 1, 1};
    // primitiveMesh::cells(): faces a cell owns (ascending) then faces where it is neighbour (ascending)
    veci cellFaceStart, cellFaces;
    vecd dCoupled;  // [3*NB] delta vector of coupled boundary faces (own delta - neighbour delta)
    bool coupled(const Patch& p) const { return p.kind == ICSB200_CYCLIC || p.kind == ICSB200_PROCESSOR || p.kind == ICSB200_CYCLICAMI; }
    bool empty(const Patch& p) const { return p.kind == ICSB200_EMPTY; }
};

struct BC {
    int kind = ICSB200_BC_ZEROGRADIENT;
    double prm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // non-uniform entries (`nonuniform List<...>` of p0, T0, value, inletValue, ...): 8 doubles per face of the patch, or empty
    vecd prmFace;
    const double* P(int faceInPatch) const { return prmFace.empty() ? prm : &prmFace[(size_t)8 * faceInPatch]; }
};

// one LDU sub-block of the coupledMatrix: blockFvMatrix<sourceType, blockType> (blockFvMatrix.H)
struct Blk {
    int nc = 1;  // doubles per coefficient: 1 scalar, 3 vector, 9 tensor
    bool exists = false, hasOff = false, hasInt = false;
    vecd diag, upper, lower, intUpper, intLower, source;
};

// rank-to-rank plumbing standing in for OpenFOAM's Pstream (one oracle context per "MPI rank")
struct Comm {
    int rank = 0, size = 1;
    virtual ~Comm() {}
    virtual double sum(double x) { return x; }
    virtual double max(double x) { return x; }
    virtual double min(double x) { return x; }
    // exchange n doubles per face with the neighbour rank of a processor patch
    virtual void exchange(int /*nbrRank*/, const double* /*send*/, double* /*recv*/, int /*n*/) {}
};

struct Ctx {
    Mesh m;
    Comm defaultComm;
    Comm* comm = &defaultComm;
    double R = 287, Cp = 1005, Cv = 718, gamma = 1.4, mu = 0, Pr = 1;
    icsb200_schemes sch{};
    std::vector<std::array<BC, 3>> bc;  // per patch: p, U, T
    bool meshSet = false, stateSet = false, matrixSet = false;

    // vol fields: cell values [0,N) then boundary values [N, N+NB) (coupled patches: patchNeighbourField)
    vecd p, T, e, psi, rho, rhoE;  // scalar
    vecd U, rhoU;                  // vector (AoS)
    // valueInternalCoeffs of p/U/T per boundary face, frozen when the BCs were last evaluated
    // (mixedFvPatchField keeps valueFraction_ from its last updateCoeffs)
    vecd vicP, vicU, vicT;
    // gradientInternalCoeffs of U / T per boundary face, frozen with them (viscousFluxScheme.C:58-59, full viscous Jacobian)
    vecd gicU, gicT;
    // old-time levels of the conserved variables (cells only)
    vecd rho0, rhoU0, rhoE0, rho00, rhoU00, rhoE00;
    int timeIndex = 0;
    // surface fields [FT]
    vecd phi, phiUp, phiEp;
    bool phiValid = false;
    // pseudo time
    vecd rPseudoDeltaT, pseudoCoField, ddtCoeff;
    double pseudoCoNum = 1;
    bool haveInitRes = false, havePrevRes = false, firstIter = true;
    icsb200_residuals initRes{}, prevRes{};
    // sources R*V
    vecd srcRho, srcRhoU, srcRhoE;
    // previous-iteration conserved variables
    vecd rhoPrev, rhoUPrev, rhoEPrev;
    // increments
    vecd dRho, dRhoU, dRhoE;
    // coupledMatrix(mesh, 2, 1): block ids as in icsb200_matrix_get_ldu
    Blk blk[9];
    // Roe dissipation members (roeFluxScheme.H:61-71) are temporaries here
    // cyclicAMI tables handed over by orc_ami_set before orc_mesh_set: patch -> (start, face, weight)
    std::vector<std::pair<int, Patch>> pendingAmi;
    // MRF (convectiveFluxScheme.H MRFFaceVelocity_/MRFOmega_, set by the solver at outerLoop.H:18-21): frame velocity
    // normal to each face [FT] (face orientation) and frame angular velocity per cell [3N]; empty = zero fields
    vecd mrfFaceVel, mrfOmega;
    // turbulence->muEff() / alphaEff() handed over by the caller (cells then boundary faces, [N+NB]); empty = laminar
    // constants mu and gamma mu / Pr.  Coupled boundary slots are filled with the patchNeighbourField.
    vecd muEffField, alphaEffField;
    // phase-lag cyclic patches (src/fields/fvPatchFields/constraint/phaseLagCyclic): this context is time instance hbIndex of
    // hbSiblings; lagRow[patch] = row hbIndex of D_pl = Re(EInv M(IBPA) E) for that patch side.  The patchNeighbourField of the
    // listed fields (U*, p, rho, E, H, c) is sum_J lagRow[J] * (field of instance J at the neighbour cell)
    std::vector<Ctx*> hbSiblings;
    int hbIndex = 0;
    std::map<int, vecd> lagRow;
    bool srcMrfApplied = false;  // the Coriolis source has been subtracted from the current srcRhoU
    double mrfAt(int f) const { return mrfFaceVel.empty() ? 0.0 : mrfFaceVel[f]; }
    std::string err;
};

}  // namespace orc
