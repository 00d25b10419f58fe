/* icsb200.h — C ABI of the B200-native ICSFoam implicit pseudo-time iteration hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / OpenFOAM / CUDA
 * types.  Every entry point names the reference interface it replaces (paths relative to the
 * ICSFoam tree).  Arrays are caller-owned HOST pointers in OpenFOAM's native AoS layout
 * (vector = 3 contiguous doubles, tensor = 9 row-major doubles, label = 32-bit int) so that
 * Field<vector>::cdata() can be passed without a copy; the device layout is private.
 * Entry points ending in _dev operate on state already resident in HBM (no host transfer).
 *
 * All functions return 0 on success or a negative ICSB200_E* code; icsb200_last_error() gives the
 * message (the OpenFOAM adapter turns non-zero into FatalErrorInFunction, mirroring
 * coupledMatrixSolver.C:53-61, lusgs.C:82-85).  One context per rank / GPU, one host thread.
 * There is no CPU fallback: icsb200_create fails if no sm_100 device is present.
 */
#ifndef ICSB200_H
#define ICSB200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct icsb200_ctx icsb200_ctx;

enum {
    ICSB200_OK = 0,
    ICSB200_EINVAL = -1,   /* bad argument / unknown selector (cf. newConvectiveFluxScheme.C:56-66) */
    ICSB200_ESTATE = -2,   /* call order violated (mesh/state/matrix not set) */
    ICSB200_ECUDA = -3,    /* CUDA / NCCL runtime failure */
    ICSB200_ESINGULAR = -4 /* "All diagonals of coupledMatrix are zero" lusgs.C:118-123 */
};

/* fvPatch kinds as the hot path distinguishes them (convectiveFluxScheme.C:243, setCoAndDeltaT.H:72-86) */
enum { ICSB200_PATCH = 0, ICSB200_WALL = 1, ICSB200_EMPTY = 2, ICSB200_SYMMETRYPLANE = 3, ICSB200_CYCLIC = 4, ICSB200_PROCESSOR = 5,
    ICSB200_CYCLICAMI = 6 /* cyclicAMI: patchNeighbourField = weighted sum over neighbour-patch faces (icsb200_ami_set) */ };

typedef struct {
    int kind;           /* ICSB200_PATCH ... */
    int start, size;    /* face range [start, start+size) in the global face list */
    int nbr_rank;       /* PROCESSOR: neighbProcNo */
    int nbr_patch;      /* CYCLIC: index of the neighbour patch */
    double forwardT[9]; /* CYCLIC: cyclicPolyPatch::forwardT() row-major — patchNeighbourField = transform(forwardT, neighbour
                         * value) (cyclicFvPatchField.C:130-190): identity = translational pair; a rotation = rotational pair
                         * (tensors grad(U), tauMC of the viscous terms: transform(forwardT, t)); CYCLICAMI: applied after the
                         * interpolation (cyclicAMIFvPatchField.C:171-203) */
} icsb200_patch;

/* run-time selectors — same words as the reference dictionaries */
enum { ICSB200_FLUX_HLLC = 0, ICSB200_FLUX_ROE = 1, ICSB200_FLUX_AUSMPLUSUP = 2,       /* fvSchemes convectiveFluxScheme/fluxScheme */
       ICSB200_FLUX_RUSANOV = 3 };  /* NOT a reference scheme (src/Make/files:47-51 has none): the local Lax-Friedrichs flux whose
                                       linearisation the reference's approximate Jacobian is (convectiveFluxScheme.C:402-546);
                                       named by the project brief, parity unpinned by construction */
enum { ICSB200_LIM_UPWIND = 0, ICSB200_LIM_VANLEER = 1, ICSB200_LIM_MINMOD = 2, ICSB200_LIM_LINEAR = 3 }; /* interpolationSchemes reconstruct(.) */
enum { ICSB200_DDT_STEADY = 0, ICSB200_DDT_EULER = 1, ICSB200_DDT_BACKWARD = 2 };       /* ddtSchemes: dualTime rPseudoDeltaT <inner> */
enum { ICSB200_SOLVER_GMRES = 0, ICSB200_SOLVER_SMOOTH = 1 };                            /* fvSolution flowSolver/solver: GMRES,
                                                                                          * smoothSolverCoupled (smoother Jacobi) */
enum { ICSB200_PRECOND_LUSGS = 0, ICSB200_PRECOND_JACOBI = 1 };                          /* flowSolver/<solver>/preconditioner */

/* boundary-condition kinds for p, U, T (OpenFOAM fvPatchField type names) */
enum {
    ICSB200_BC_ZEROGRADIENT = 0,
    ICSB200_BC_FIXEDVALUE = 1,                  /* params: value (1 or 3 doubles) */
    ICSB200_BC_SLIP = 2,                        /* also symmetryPlane (basicSymmetry) */
    ICSB200_BC_EMPTY = 3,
    ICSB200_BC_INLETOUTLET = 4,                 /* params: inletValue; also 'freestream' (mixed on phi) */
    ICSB200_BC_TOTALPRESSURE = 5,               /* params: p0, gamma */
    ICSB200_BC_TOTALTEMPERATURE = 6,            /* params: T0, gamma */
    ICSB200_BC_PRESSUREINLETOUTLETVELOCITY = 7, /* params: tangentialVelocity (3) */
    ICSB200_BC_FREESTREAMPRESSURE = 8,          /* params: freestreamValue, Uinf (3) */
    ICSB200_BC_COUPLED = 9                      /* cyclic / processor: set automatically */
};
enum { ICSB200_FIELD_P = 0, ICSB200_FIELD_U = 1, ICSB200_FIELD_T = 2 };

typedef struct {
    int flux_scheme;          /* ICSB200_FLUX_*        (newConvectiveFluxScheme.C:50) */
    int limiter_rho;          /* reconstruct(rho): rho, p  (hllcFluxScheme.C:84-88) */
    int limiter_U;            /* reconstruct(U)            (hllcFluxScheme.C:92-97) */
    int limiter_T;            /* reconstruct(T): c, E, H   (hllcFluxScheme.C:105-121) */
    int low_mach_ausm;        /* lowMachAusm, default 1    (ausmPlusUpFluxScheme.C:61) */
    double entropy_fix_coeff; /* entropyFixCoeff, default 0.05 (roeFluxScheme.C:255) */
    int ddt_scheme;           /* ICSB200_DDT_*  inner scheme of 'dualTime rPseudoDeltaT <inner>' (dualTimeDdtScheme.H:109-117) */
    double delta_t;           /* physical time step (transient) */
    int local_timestepping;   /* pseudoTime/localTimestepping, default 1 (initialise.H:39-43) */
    int local_timestepping_bounding; /* default 1 (initialise.H:52-57) */
    double local_timestepping_lower_bound; /* default 0.95 (initialise.H:59-69) */
    double pseudo_co_num;     /* pseudoCoNum (createFields.H:218-254) */
    double pseudo_co_num_min, pseudo_co_num_max;          /* beginTimeStep.H:29-33 */
    double pseudo_co_num_max_incr, pseudo_co_num_min_decr; /* beginTimeStep.H:35-45 */
    double rho_min, T_min, T_max;                          /* updateFields.H:11-35 */
    int viscous_full_jacobian; /* 0 (default): viscousFluxScheme/LaxFriedrichJacobian true (viscousFluxScheme.C:228-240);
                                * 1: LaxFriedrichJacobian false — the fvj::laplacian blocks of :248-261 and the wall terms of
                                * viscousFluxScheme::addBoundaryTerms (:120-215, gradientInternalCoeffs) */
} icsb200_schemes;

typedef struct {
    int solver;         /* ICSB200_SOLVER_*  (coupledMatrixSolver.C:41-67; gmres.C:772-1110, smoothSolverCoupled.C:385-515) */
    int preconditioner; /* ICSB200_PRECOND_*   (coupledMatrixPreconditioner.C:41-64) */
    int n_directions;   /* GMRES: nDirections (gmres.C:81); smoothSolverCoupled: nSweeps (smoothSolverCoupled.C:57), the block-Jacobi
                         * sweeps of JacobiSmoother::smooth (JacobiSmoother.C:120-203) between two residual evaluations */
    int max_iter;       /* maxIter, default 1000 (coupledMatrixSolver.C:81, coupledMatrix.C:40) */
    int min_iter;       /* minIter, default 0 */
    double tolerance;   /* tolerance */
    double rel_tol;     /* relTol */
} icsb200_solver_controls;

/* residualsIO (residualsIO.H:54-312) for nScalar=2 (rho, rhoE), nVector=1 (rhoU) */
typedef struct {
    double s_init[2], v_init[3];
    double s_final[2], v_final[3];
    int n_iterations;
} icsb200_residuals;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* nccl_unique_id: 128-byte ncclUniqueId shared by all ranks (NULL when n_ranks == 1). */
int icsb200_create(icsb200_ctx** ctx, int device, const void* nccl_unique_id, int rank, int n_ranks);
int icsb200_destroy(icsb200_ctx* ctx);
const char* icsb200_last_error(icsb200_ctx* ctx);
/* fills a 128-byte buffer with a fresh ncclUniqueId (rank 0 calls this, then broadcasts it) */
int icsb200_nccl_unique_id(void* out128);

/* ---- setup --------------------------------------------------------------------------------- */
/* fvMesh as the reference reads it: mesh.owner()/neighbour() (lduAddr, lusgs.C:141-156), Sf/magSf
 * (hllcFluxScheme.C:78), weights (convectiveFluxScheme.C:376), deltaCoeffs/nonOrthDeltaCoeffs
 * (setCoAndDeltaT.H:61,145), C (limited schemes), V (residualsUpdate.H:81-83), Cf (coupled patch
 * deltas), boundary patches, solutionD (coupledMatrix.C:371-382).  n_faces includes boundary faces. */
int icsb200_mesh_set(icsb200_ctx* ctx, int n_cells, int n_internal_faces, int n_faces, const int* owner,
                     const int* neighbour, const double* Sf, const double* magSf, const double* weights,
                     const double* deltaCoeffs, const double* nonOrthDeltaCoeffs, const double* C, const double* V,
                     const double* Cf, int n_patches, const icsb200_patch* patches, const int solutionD[3]);
/* cyclicAMI interpolation of patch `patch` (kind ICSB200_CYCLICAMI, neighbour = nbr_patch): for face i of the patch,
 * patchNeighbourField[i] = sum_{k in [face_start[i], face_start[i+1])} weight[k] * phi[faceCells(nbr_patch)[nbr_face[k]]]
 * in this order (AMIInterpolation::interpolateToSource/Target with plusEqOp; originalOFFiles/constraintFvPatchFields/
 * cyclicAMI/cyclicAMIFvPatchField.C:146-209, no lowWeightCorrection; a rotational pair then applies transform(forwardT, .)).  The addressing
 * and weights are OpenFOAM's (cyclicAMIPolyPatch::AMI().srcAddress()/srcWeights()); the mesh arrays weights /
 * deltaCoeffs / nonOrthDeltaCoeffs of the patch faces are cyclicAMIFvPatch's.  Call BEFORE icsb200_mesh_set, once per
 * cyclicAMI patch; both patches of a pair live on the same rank. */
int icsb200_ami_set(icsb200_ctx* ctx, int patch, int n_faces, const int* face_start, const int* nbr_face, const double* weight);
/* phaseLagCyclic pair (src/fields/fvPatchFields/constraint/phaseLagCyclic/phaseLagCyclicFvPatchField.C:160-398) in a Harmonic
 * Balance run: `patch` is a CYCLIC patch of time instance K of the instance-replicated mesh (see icsb200_hb_set) and
 * weights[n_instants] is row K of D_pl = Re(EInv M(IBPA) E) for that side of the pair (owner +IBPA, neighbour -IBPA; host set-up,
 * icsfoam_b200/hb.py phase_lag_operator).  The patchNeighbourField of the fields the reference lists by name (rho, p, U, E, H, c) is
 * then sum_J weights[J] * (value of instance J at the neighbour cell), followed by transform(forwardT, .) on rotational pairs;
 * all other fields (gradients, solver operands, T) keep the plain cyclic value.  Call BEFORE icsb200_mesh_set, once per replicated
 * patch; n_instants <= 16. */
int icsb200_phaselag_set(icsb200_ctx* ctx, int patch, int n_instants, const double* weights);
/* hePsiThermo<pureMixture<constTransport<hConst<perfectGas>>>>, sensibleInternalEnergy (createFields.H:17-35).
 * mu > 0 makes the run viscous (createFields.H:37-45 `inviscid`): icsb200_residual / iterate then add the laminar
 * viscous terms of residualsUpdate.H:16-43 (laplacian(muEff,U), div(tauMC), div(sigmaDotU & Sf), laplacian(alphaEff,e),
 * `Gauss linear corrected`) and icsb200_assemble the viscous Jacobian — the Lax-Friedrichs branch (viscousFluxScheme.C:220-246)
 * or, with icsb200_schemes::viscous_full_jacobian, the full one (:248-261 + addBoundaryTerms :120-215). */
int icsb200_thermo_set(icsb200_ctx* ctx, double R, double Cp, double mu, double Pr);
int icsb200_schemes_set(icsb200_ctx* ctx, const icsb200_schemes* s);
/* fvPatchField of p, U or T on one patch (0/p, 0/U, 0/T boundaryField entries) */
int icsb200_bc_set(icsb200_ctx* ctx, int patch, int field, int kind, const double* params, int n_params);
/* the same with non-uniform entries (`nonuniform List<...>` of value, p0, T0, inletValue, tangentialVelocity ...): params is
 * [patch size][n_params], one row per face of the patch in its face order — inlet profiles, radially varying total pressure */
int icsb200_bc_set_nonuniform(icsb200_ctx* ctx, int patch, int field, int kind, const double* params, int n_params);

/* Multiple reference frames: the two fields the solver hands to the flux scheme every outer iteration
 * (applications/solvers/dbnsFoam/outerLoop.H:18-21): mrf_face_velocity[n_faces] = flux.MRFFaceVelocity() =
 * (MRF.faceU() + MRFTrans.faceU()) & Sf/magSf, in each face's own orientation (boundary faces outward), and
 * mrf_omega[3*n_cells] = flux.MRFOmega().  NULL = zero field (the default).  They enter the three flux schemes
 * (hllcFluxScheme.C:157-161,217-218; roeFluxScheme.C:362-363,404-408; ausmPlusUpFluxScheme.C:101-102,293), the spectral
 * radius of the pseudo time step and of the dissipation Jacobian (setCoAndDeltaT.H:39-55,97-125;
 * convectiveFluxScheme.C:498-524), the boundary Jacobian (convectiveFluxScheme.C:94), the fvj::div(w, MRFFaceVelocity
 * magSf) blocks (convectiveFluxScheme.C:477-481; blockFvOperatorsTemplates.C:146-201) and addMRFSource
 * (convectiveFluxScheme.C:123-139: Coriolis source and the skew momentum-diagonal entries).  MRF zone set-up
 * (MRFCoupledZone faceU/omega, correctBoundaryVelocity) stays on the host.  Call after icsb200_mesh_set. */
int icsb200_mrf_set(icsb200_ctx* ctx, const double* mrf_face_velocity, const double* mrf_omega);

/* Effective transport properties from the caller's turbulence model: turbulence->muEff() and turbulence->alphaEff() as
 * viscousFluxScheme::addFluxTerms (viscousFluxScheme.C:222-223) and residualsUpdate.H:16-43 read them every outer
 * iteration — cell values [n_cells] and boundary-face values [n_faces - n_internal_faces] (wall functions etc.; entries of
 * coupled and empty patches are ignored).  The turbulence transport equations themselves (OpenFOAM's
 * compressible::turbulenceModel, dbnsFoam.C:126-134) stay with the caller.  NULL muEff = laminar constants mu and
 * gamma mu / Pr from icsb200_thermo_set (the default).  Needs mu > 0 (a viscous run).  Call after icsb200_mesh_set. */
int icsb200_transport_set(icsb200_ctx* ctx, const double* muEff, const double* muEff_boundary, const double* alphaEff,
                          const double* alphaEff_boundary);

/* ---- state --------------------------------------------------------------------------------- */
/* internal fields p[N], U[3N], T[N]; evaluates BCs + thermo and builds rho, rhoU, rhoE (createFields.H:75-131) */
int icsb200_state_set(icsb200_ctx* ctx, const double* p, const double* U, const double* T);
/* any pointer may be NULL.  Cell arrays [N]/[3N]. */
int icsb200_state_get(icsb200_ctx* ctx, double* rho, double* rhoU, double* rhoE, double* p, double* U, double* T);
/* boundary values, indexed by (face - n_internal_faces); any pointer may be NULL.  Faces of coupled patches report the
 * patchNeighbourField (neighbour cell value: interpolated for cyclicAMI, rotated on rotational pairs, phase-lagged where that
 * applies).  With processor patches this call exchanges a halo, i.e. it is COLLECTIVE over the ranks (as are icsb200_state_set,
 * icsb200_transport_set and every iterate / matrix_mul / solve call). */
int icsb200_boundary_get(icsb200_ctx* ctx, double* rho_b, double* U_b, double* p_b, double* T_b);
/* transient: shift W -> W.old -> W.oldOld (runTime++ of dbnsFoam.C:103) */
int icsb200_new_time_step(icsb200_ctx* ctx);

/* ---- the hot path, piecewise (for parity and for partial offload) ---------------------------- */
/* convectiveFluxScheme::calcFlux(phi, phiUp, phiEp) (convectiveFluxScheme.H:219-224;
 * hllcFluxScheme.C:70-240, roeFluxScheme.C:276-409, ausmPlusUpFluxScheme.C:73-299).
 * Outputs sized n_faces / 3 n_faces / n_faces (reference face order); may be NULL to keep on device. */
int icsb200_calc_flux(icsb200_ctx* ctx, double* phi, double* phiUp, double* phiEp);
/* residualsUpdate.H:1-83 — sources R*V of the three equations: rhoR[N], rhoUR[3N], rhoER[N] (may be NULL) */
int icsb200_residual(icsb200_ctx* ctx, double* rhoR, double* rhoUR, double* rhoER);
/* setCoAndDeltaT.H:1-173 (SER + local pseudo time step); outputs rPseudoDeltaT[N], pseudoCoField[N] (may be NULL) */
int icsb200_pseudo_dt(icsb200_ctx* ctx, double* rPseudoDeltaT, double* pseudoCo);
/* convectiveFluxScheme::createConvectiveJacobian (convectiveFluxScheme.C:537-546) [+ viscousFluxScheme::
 * createViscousJacobian LF branch (viscousFluxScheme.C:220-246) when mu > 0] into device block storage.
 * ddtCoeff as outerLoop.H:61-64 is computed internally from the current rPseudoDeltaT. */
int icsb200_assemble(icsb200_ctx* ctx);
/* read back one LDU sub-block of the coupledMatrix (coupledMatrix.H:399-421) in the reference layout.
 * block: 0 dSByS(0,0) 1 dSByS(0,1) 2 dSByS(1,0) 3 dSByS(1,1) 4 dSByV(0,0) 5 dSByV(1,0) 6 dVByS(0,0)
 * 7 dVByS(0,1) 8 dVByV(0,0).  diag[nc*N], upper[nc*F], lower[nc*F] with nc = 1, 3 or 9; NULL skips. */
int icsb200_matrix_get_ldu(icsb200_ctx* ctx, int block, double* diag, double* upper, double* lower);
/* accept a HOST-assembled coupledMatrix (solver-only drop-in); NULL upper/lower = no off-diagonal */
int icsb200_matrix_set_ldu(icsb200_ctx* ctx, int block, const double* diag, const double* upper, const double* lower);
/* interfacesUpper() of one sub-block on the coupled patches (processor, cyclic, cyclicAMI), the coefficients Amul applies to
 * the patchNeighbourField (blockFvMatrix.C:248-266, 383-599): intUpper[nc * (n_faces - n_internal_faces)] in boundary-face
 * order; entries of non-coupled faces are written as 0 by get and ignored by set.  A solver-only adapter on a decomposed
 * or periodic case calls set after icsb200_matrix_set_ldu for every sub-block that has interfaces. */
int icsb200_matrix_get_interfaces(icsb200_ctx* ctx, int block, double* intUpper);
int icsb200_matrix_set_interfaces(icsb200_ctx* ctx, int block, const double* intUpper);
/* sources of the three equations (dSByS(0,0), dVByV(0,0), dSByS(1,1) .source(); residualsUpdate.H:81-83) */
int icsb200_source_set(icsb200_ctx* ctx, const double* sRho, const double* sRhoU, const double* sRhoE);
/* the sources the solver sees: R*V (+ HB source, + the MRF Coriolis term after icsb200_assemble); NULL skips */
int icsb200_source_get(icsb200_ctx* ctx, double* sRho, double* sRhoU, double* sRhoE);
/* coupledMatrix::matrixMul (coupledMatrix.C:66-123): y = A x for x = (rho[N], rhoU[3N], rhoE[N]) */
int icsb200_matrix_mul(icsb200_ctx* ctx, const double* xRho, const double* xRhoU, const double* xRhoE, double* yRho,
                       double* yRhoU, double* yRhoE);
/* coupledMatrix::preconditioner::precondition in place (coupledMatrix.H:363-367; lusgs.C:220-382, Jacobi.C:55-132) */
int icsb200_precondition(icsb200_ctx* ctx, int preconditioner, double* xRho, double* xRhoU, double* xRhoE);
/* coupledMatrix::solveForIncr -> gmres::solveDelta 6-arg (coupledMatrix.C:297-385, gmres.C:772-1110).
 * W = current conserved state on device; outputs the increment (may be NULL to keep on device). */
int icsb200_solve_delta(icsb200_ctx* ctx, const icsb200_solver_controls* c, double* dRho, double* dRhoU, double* dRhoE,
                        icsb200_residuals* res);
/* boundLocalTimeStep.H:1-98 + updateFields.H:1-104 */
int icsb200_update_fields(icsb200_ctx* ctx);

/* ---- the hot path, fused: one outer pseudo-time iteration of dbnsFoam (outerLoop.H:51-99 + updateFields.H) -- */
/* device-resident: flux -> residual -> pseudo dt -> Jacobian -> GMRES/LU-SGS -> update (outerLoop.H:51-99, updateFields.H:1-104) */
int icsb200_iterate_dev(icsb200_ctx* ctx, const icsb200_solver_controls* c, icsb200_residuals* res);
/* same through host buffers: uploads p,U,T (6N doubles), iterates once, downloads p,U,T — the call an
 * OpenFOAM adapter makes when fields live on the host */
int icsb200_iterate_host(icsb200_ctx* ctx, const icsb200_solver_controls* c, double* p, double* U, double* T,
                         icsb200_residuals* res);

/* ---- Harmonic Balance (dbnsFullyImplicitHBFoam) ------------------------------------ */
/* HBZone / HBZoneList (src/cfdTools/HB/HBZone.C:270-356 operators, :435-518 addBlock, :521-651 cylindrical source;
 * HBZoneTemplates.C:38-92 addSource) and the (2 nO, nO) coupled system of
 * applications/solvers/dbnsFullyImplicitHBFoam/outerLoop.H:28-30,91-206.
 * The n_instants time instances ("subTimeLevelK" meshes) are passed to icsb200_mesh_set as ONE mesh made of n_instants
 * disconnected copies, instance-major in cells, faces and patches (cell c of instance K is cell K*N/n_instants + c; host
 * tooling: meshtools replicate).  All field arrays of the other entry points are then instance-major as well.
 * D[z] (n_instants x n_instants, row-major) is HBZone::D() of zone z; zone_of_cell[N/n_instants] gives the zone of every
 * instance cell (-1: none; NULL: every cell is in zone 0 = `allMesh`).  cyl_coords[z] != 0 selects the cylindrical
 * momentum source with rotation_axis[3z..] / rotation_centre[3z..].  Call after icsb200_mesh_set.
 * Effect: residual adds S_J = -V sum_K D[J][K] W_K; assemble adds V D[J][J] to the (rho,rho), (rhoU,rhoU), (rhoE,rhoE)
 * diagonals; matrix_mul / solve_delta include the inter-instance diagonal coupling V D[J][K]; LU-SGS uses the shared
 * rDiagCoeff over all instances (lusgs.C:50-123); the SER ratio uses the residual norm over all instances
 * (outerLoop.H:32-64); residuals are per instance (icsb200_hb_residuals_get). n_instants = 1 switches HB off.
 * Multi-rank: every rank passes the n_instants copies of ITS partition (processor patches replicated per instance,
 * instance-major like all other patches); instances are never split across ranks. */
int icsb200_hb_set(icsb200_ctx* ctx, int n_instants, int n_zones, const double* D, const int* zone_of_cell,
                   const int* cyl_coords, const double* rotation_axis, const double* rotation_centre);
/* residualsIO of the last solve for all instances: s_* [2 n_instants] = (rho_0, rhoE_0, rho_1, ...), v_* [3 n_instants]
 * (residualsIO.H:204-240; dbnsFullyImplicitHBFoam/setUpResiduals.H:1-11) */
int icsb200_hb_residuals_get(icsb200_ctx* ctx, double* s_init, double* v_init, double* s_final, double* v_final);

/* ---- introspection ------------------------------------------------------------------------- */
/* number of kernels this library launched since create (bench.py "gpu_launches") */
long long icsb200_launch_count(icsb200_ctx* ctx);
/* per-kernel-class device time in ms accumulated since the last reset (CUDA events on the compute stream).
 * names: NUL-separated list written to names_buf; returns the number of classes */
int icsb200_timers_get(icsb200_ctx* ctx, char* names_buf, int names_len, double* ms, long long* calls, int max_classes);
int icsb200_timers_reset(icsb200_ctx* ctx, int enable);
/* device-side stopwatch: CUDA events recorded on the library's compute stream (bench.py times its steps with these) */
int icsb200_timer_begin(icsb200_ctx* ctx);
int icsb200_timer_end(icsb200_ctx* ctx, double* elapsed_ms);
/* LU-SGS schedule statistics: n_levels_fwd, n_levels_rev, max_level_width, n_positions, tile_mode (0/1), n_tiles,
 * n_tile_levels, reserved */
int icsb200_schedule_info(icsb200_ctx* ctx, int out[8]);

#ifdef __cplusplus
}
#endif
#endif
