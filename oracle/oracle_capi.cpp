// oracle_capi.cpp — C entry points of the CPU oracle, mirroring include/icsb200.h one for one (prefix orc_)
// so that the parity tests can drive the oracle and the CUDA product with the same script.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
#include <condition_variable>
#include <mutex>
#include <thread>

#include "oracle_internal.hpp"

using namespace orc;

namespace {

// ---- thread "ranks": a World of P oracle contexts run SPMD by P threads; stands in for MPI (Pstream) ----
struct World {
    int size;
    std::mutex mu;
    std::condition_variable cv;
    int count = 0;
    long long gen = 0;
    std::vector<double> red;
    std::vector<const double*> box;  // box[src*size + dst] = send buffer
    explicit World(int n) : size(n), red(n, 0.0), box((size_t)n * n, nullptr) {}
    void barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        long long g = gen;
        if (++count == size) { count = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

struct ThreadComm : Comm {
    World* w;
    ThreadComm(World* w_, int r) : w(w_) { rank = r; size = w_->size; }
    template <class Op> double reduce(double x, Op op)
    {
        w->red[rank] = x;
        w->barrier();
        double r = w->red[0];
        for (int i = 1; i < size; i++) r = op(r, w->red[i]);  // rank order: deterministic
        w->barrier();
        return r;
    }
    double sum(double x) override { return reduce(x, [](double a, double b) { return a + b; }); }
    double max(double x) override { return reduce(x, [](double a, double b) { return a > b ? a : b; }); }
    double min(double x) override { return reduce(x, [](double a, double b) { return a < b ? a : b; }); }
    void exchange(int nbr, const double* send, double* recv, int n) override
    {
        // every rank visits its processor patches in ascending neighbour order; pairwise mailboxes make
        // the exchange independent of that order
        w->box[(size_t)rank * size + nbr] = send;
        // wait until the neighbour has posted for us
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.notify_all();
            w->cv.wait(lk, [&] { return w->box[(size_t)nbr * size + rank] != nullptr; });
        }
        std::memcpy(recv, w->box[(size_t)nbr * size + rank], sizeof(double) * n);
        // handshake: tell the neighbour we are done with its buffer, wait for it to be done with ours
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->box[(size_t)nbr * size + rank] = nullptr;
            w->cv.notify_all();
            w->cv.wait(lk, [&] { return w->box[(size_t)rank * size + nbr] == nullptr; });
        }
    }
};

int fail(Ctx* c, int code, const char* msg) { c->err = msg; return code; }

}  // namespace

extern "C" {

int orc_create(Ctx** out) { *out = new Ctx; return 0; }
int orc_destroy(Ctx* c) { if (c->comm != &c->defaultComm) delete c->comm; delete c; return 0; }
const char* orc_last_error(Ctx* c) { return c->err.c_str(); }

// a World groups n contexts that will be driven by n threads (see orc_world_run)
void* orc_world_create(int n) { return new World(n); }
void orc_world_destroy(void* w) { delete (World*)w; }
int orc_attach(Ctx* c, void* world, int rank) { c->comm = new ThreadComm((World*)world, rank); return 0; }

int orc_mesh_set(Ctx* c, int n_cells, int n_internal_faces, int n_faces, const int* owner, const int* neighbour, const double* Sf,
                 const double* magSf, const double* weights, const double* deltaCoeffs, const double* nonOrthDeltaCoeffs, const double* C,
                 const double* V, const double* Cf, int n_patches, const icsb200_patch* patches, const int solutionD[3])
{
    Mesh& m = c->m;
    m.N = n_cells; m.F = n_internal_faces; m.FT = n_faces;
    m.owner.assign(owner, owner + n_faces);
    m.neighbour.assign(neighbour, neighbour + n_internal_faces);
    m.Sf.assign(Sf, Sf + 3 * (size_t)n_faces);
    m.magSf.assign(magSf, magSf + n_faces);
    m.w.assign(weights, weights + n_faces);
    m.deltaCoeffs.assign(deltaCoeffs, deltaCoeffs + n_faces);
    m.nonOrthDeltaCoeffs.assign(nonOrthDeltaCoeffs, nonOrthDeltaCoeffs + n_faces);
    m.C.assign(C, C + 3 * (size_t)n_cells);
    m.V.assign(V, V + n_cells);
    if (Cf) m.Cf.assign(Cf, Cf + 3 * (size_t)n_faces); else m.Cf.assign(3 * (size_t)n_faces, 0.0);
    m.patches.clear();
    for (int i = 0; i < n_patches; i++) {
        Patch p;
        p.kind = patches[i].kind; p.start = patches[i].start; p.size = patches[i].size;
        p.nbrRank = patches[i].nbr_rank; p.nbrPatch = patches[i].nbr_patch;
        std::memcpy(p.forwardT, patches[i].forwardT, sizeof(p.forwardT));
        if (p.kind == ICSB200_CYCLICAMI) {
            bool found = false;
            for (auto& pa : c->pendingAmi)
                if (pa.first == i) { p.amiStart = pa.second.amiStart; p.amiFace = pa.second.amiFace; p.amiWeight = pa.second.amiWeight; found = true; }
            if (!found || (int)p.amiStart.size() != p.size + 1) return fail(c, ICSB200_EINVAL, "cyclicAMI patch without a matching orc_ami_set");
        }
        if (p.kind == ICSB200_CYCLIC || p.kind == ICSB200_CYCLICAMI)
            for (int k = 0; k < 9; k++)
                if (std::fabs(p.forwardT[k] - (k % 4 == 0 ? 1.0 : 0.0)) > 1e-12) p.rotational = true;
        m.patches.push_back(p);
    }
    for (int d = 0; d < 3; d++) m.solutionD[d] = solutionD[d];
    for (int f = 0; f < m.F; f++) if (m.owner[f] >= m.neighbour[f]) return fail(c, ICSB200_EINVAL, "mesh is not in upper-triangular order");
    meshFinalize(*c);
    c->meshSet = true;
    return 0;
}

int orc_ami_set(Ctx* c, int patch, int n_faces, const int* face_start, const int* nbr_face, const double* weight)
{
    Patch p{};
    p.amiStart.assign(face_start, face_start + n_faces + 1);
    p.amiFace.assign(nbr_face, nbr_face + face_start[n_faces]);
    p.amiWeight.assign(weight, weight + face_start[n_faces]);
    for (auto& pa : c->pendingAmi) if (pa.first == patch) { pa.second = p; return 0; }
    c->pendingAmi.emplace_back(patch, p);
    return 0;
}

int orc_thermo_set(Ctx* c, double R, double Cp, double mu, double Pr)
{
    c->R = R; c->Cp = Cp; c->Cv = Cp - R; c->gamma = Cp / c->Cv; c->mu = mu; c->Pr = Pr;
    return 0;
}

int orc_schemes_set(Ctx* c, const icsb200_schemes* s)
{
    if (s->flux_scheme < 0 || s->flux_scheme > ICSB200_FLUX_RUSANOV) return fail(c, ICSB200_EINVAL, "Unknown convectiveFluxScheme type");
    c->sch = *s;
    return 0;
}

int orc_bc_set(Ctx* c, int patch, int field, int kind, const double* params, int n_params)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    if (patch < 0 || patch >= (int)c->bc.size() || field < 0 || field > 2 || n_params > 8) return fail(c, ICSB200_EINVAL, "bad bc");
    c->bc[patch][field].kind = kind;
    c->bc[patch][field].prmFace.clear();
    for (int i = 0; i < n_params; i++) c->bc[patch][field].prm[i] = params[i];
    return 0;
}

// non-uniform entries: params[size of the patch][n_params], one row per face (nonuniform List<...> in 0/p, 0/U, 0/T)
int orc_bc_set_nonuniform(Ctx* c, int patch, int field, int kind, const double* params, int n_params)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    if (patch < 0 || patch >= (int)c->bc.size() || field < 0 || field > 2 || n_params < 1 || n_params > 8 || !params) return fail(c, ICSB200_EINVAL, "bad bc");
    BC& b = c->bc[patch][field];
    const int n = c->m.patches[patch].size;
    b.kind = kind;
    b.prmFace.assign((size_t)8 * n, 0.0);
    for (int i = 0; i < n; i++)
        for (int k = 0; k < n_params; k++) b.prmFace[(size_t)8 * i + k] = params[(size_t)n_params * i + k];
    return 0;
}

int orc_state_set(Ctx* c, const double* p, const double* U, const double* T)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    stateInit(*c, p, U, T);
    return 0;
}

int orc_state_get(Ctx* c, double* rho, double* rhoU, double* rhoE, double* p, double* U, double* T)
{
    const int N = c->m.N;
    if (rho) std::memcpy(rho, c->rho.data(), sizeof(double) * N);
    if (rhoU) std::memcpy(rhoU, c->rhoU.data(), sizeof(double) * 3 * N);
    if (rhoE) std::memcpy(rhoE, c->rhoE.data(), sizeof(double) * N);
    if (p) std::memcpy(p, c->p.data(), sizeof(double) * N);
    if (U) std::memcpy(U, c->U.data(), sizeof(double) * 3 * N);
    if (T) std::memcpy(T, c->T.data(), sizeof(double) * N);
    return 0;
}

int orc_boundary_get(Ctx* c, double* rho_b, double* U_b, double* p_b, double* T_b)
{
    const int N = c->m.N, NB = c->m.NB;
    if (rho_b) std::memcpy(rho_b, c->rho.data() + N, sizeof(double) * NB);
    if (U_b) std::memcpy(U_b, c->U.data() + 3 * (size_t)N, sizeof(double) * 3 * NB);
    if (p_b) std::memcpy(p_b, c->p.data() + N, sizeof(double) * NB);
    if (T_b) std::memcpy(T_b, c->T.data() + N, sizeof(double) * NB);
    return 0;
}

int orc_new_time_step(Ctx* c) { newTimeStep(*c); return 0; }

int orc_calc_flux(Ctx* c, double* phi, double* phiUp, double* phiEp)
{
    if (!c->stateSet) return fail(c, ICSB200_ESTATE, "state not set");
    calcFlux(*c);
    const int FT = c->m.FT;
    if (phi) std::memcpy(phi, c->phi.data(), sizeof(double) * FT);
    if (phiUp) std::memcpy(phiUp, c->phiUp.data(), sizeof(double) * 3 * FT);
    if (phiEp) std::memcpy(phiEp, c->phiEp.data(), sizeof(double) * FT);
    return 0;
}

int orc_residual(Ctx* c, double* rhoR, double* rhoUR, double* rhoER)
{
    if (!c->phiValid) return fail(c, ICSB200_ESTATE, "calc_flux first");
    residualsUpdate(*c);
    const int N = c->m.N;
    if (rhoR) std::memcpy(rhoR, c->srcRho.data(), sizeof(double) * N);
    if (rhoUR) std::memcpy(rhoUR, c->srcRhoU.data(), sizeof(double) * 3 * N);
    if (rhoER) std::memcpy(rhoER, c->srcRhoE.data(), sizeof(double) * N);
    return 0;
}

int orc_pseudo_dt(Ctx* c, double* rPseudoDeltaT, double* pseudoCo)
{
    if (!c->stateSet) return fail(c, ICSB200_ESTATE, "state not set");
    setCoAndDeltaT(*c);
    const int N = c->m.N;
    if (rPseudoDeltaT) std::memcpy(rPseudoDeltaT, c->rPseudoDeltaT.data(), sizeof(double) * N);
    if (pseudoCo) std::memcpy(pseudoCo, c->pseudoCoField.data(), sizeof(double) * N);
    return 0;
}

int orc_assemble(Ctx* c)
{
    if (!c->stateSet) return fail(c, ICSB200_ESTATE, "state not set");
    if (c->srcRho.empty()) { c->srcRho.assign(c->m.N, 0); c->srcRhoU.assign(3 * (size_t)c->m.N, 0); c->srcRhoE.assign(c->m.N, 0); }
    computeDdtCoeff(*c);
    c->rhoPrev = c->rho; c->rhoUPrev = c->rhoU; c->rhoEPrev = c->rhoE;
    createJacobian(*c);
    return 0;
}

int orc_matrix_get_ldu(Ctx* c, int block, double* diag, double* upper, double* lower)
{
    if (!c->matrixSet) return fail(c, ICSB200_ESTATE, "matrix not assembled");
    const Blk& b = c->blk[block];
    const Mesh& m = c->m;
    if (diag) std::memcpy(diag, b.diag.data(), sizeof(double) * b.nc * m.N);
    if (upper) { if (b.hasOff) std::memcpy(upper, b.upper.data(), sizeof(double) * b.nc * m.F); else std::memset(upper, 0, sizeof(double) * b.nc * m.F); }
    if (lower) { if (b.hasOff) std::memcpy(lower, b.lower.data(), sizeof(double) * b.nc * m.F); else std::memset(lower, 0, sizeof(double) * b.nc * m.F); }
    return 0;
}

int orc_matrix_set_ldu(Ctx* c, int block, const double* diag, const double* upper, const double* lower)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    static const int ncs[9] = {1, 1, 1, 1, 3, 3, 3, 3, 9};
    Blk& b = c->blk[block];
    const Mesh& m = c->m;
    b.nc = ncs[block]; b.exists = true; b.hasInt = false;
    b.diag.assign(diag, diag + (size_t)b.nc * m.N);
    b.hasOff = upper && lower;
    if (b.hasOff) { b.upper.assign(upper, upper + (size_t)b.nc * m.F); b.lower.assign(lower, lower + (size_t)b.nc * m.F); }
    c->matrixSet = true;
    return 0;
}

int orc_matrix_get_interfaces(Ctx* c, int block, double* intUpper)
{
    if (!c->matrixSet) return fail(c, ICSB200_ESTATE, "matrix not assembled");
    const Blk& b = c->blk[block];
    const Mesh& m = c->m;
    std::memset(intUpper, 0, sizeof(double) * b.nc * m.NB);
    if (!b.hasInt) return 0;
    for (auto& p : m.patches)
        if (m.coupled(p))
            for (int f = p.start; f < p.start + p.size; f++)
                for (int k = 0; k < b.nc; k++) intUpper[(size_t)b.nc * (f - m.F) + k] = b.intUpper[(size_t)b.nc * (f - m.F) + k];
    return 0;
}

int orc_matrix_set_interfaces(Ctx* c, int block, const double* intUpper)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    Blk& b = c->blk[block];
    const Mesh& m = c->m;
    b.intUpper.assign(intUpper, intUpper + (size_t)b.nc * m.NB);
    b.intLower.assign((size_t)b.nc * m.NB, 0.0);   // only enters the diagonal at assembly (negSumDiag), not the product
    b.hasInt = true;
    return 0;
}

int orc_source_set(Ctx* c, const double* sRho, const double* sRhoU, const double* sRhoE)
{
    const int N = c->m.N;
    c->srcRho.assign(sRho, sRho + N); c->srcRhoU.assign(sRhoU, sRhoU + 3 * (size_t)N); c->srcRhoE.assign(sRhoE, sRhoE + N);
    c->srcMrfApplied = true;  // caller-provided sources are final (they already carry addMRFSource's term)
    return 0;
}

int orc_source_get(Ctx* c, double* sRho, double* sRhoU, double* sRhoE)
{
    const int N = c->m.N;
    if ((int)c->srcRho.size() != N) return fail(c, ICSB200_ESTATE, "no sources yet");
    if (sRho) std::memcpy(sRho, c->srcRho.data(), sizeof(double) * N);
    if (sRhoU) std::memcpy(sRhoU, c->srcRhoU.data(), sizeof(double) * 3 * N);
    if (sRhoE) std::memcpy(sRhoE, c->srcRhoE.data(), sizeof(double) * N);
    return 0;
}

// turbulence->muEff() / alphaEff(): cell values [n_cells] and boundary-face values [n_faces - n_internal_faces]; NULL = laminar
int orc_transport_set(Ctx* c, const double* muEff, const double* muEff_b, const double* alphaEff, const double* alphaEff_b)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    if (!muEff) { c->muEffField.clear(); c->alphaEffField.clear(); return 0; }
    if (!muEff_b || !alphaEff || !alphaEff_b) return fail(c, ICSB200_EINVAL, "transport_set: all four arrays or none");
    const int N = c->m.N, NB = c->m.NB;
    c->muEffField.assign(muEff, muEff + N); c->muEffField.insert(c->muEffField.end(), muEff_b, muEff_b + NB);
    c->alphaEffField.assign(alphaEff, alphaEff + N); c->alphaEffField.insert(c->alphaEffField.end(), alphaEff_b, alphaEff_b + NB);
    return 0;
}

// flux.MRFFaceVelocity() [n_faces] and flux.MRFOmega() [3*n_cells] (outerLoop.H:18-21); NULL = zero field
int orc_mrf_set(Ctx* c, const double* mrf_face_velocity, const double* mrf_omega)
{
    if (!c->meshSet) return fail(c, ICSB200_ESTATE, "mesh not set");
    if (mrf_face_velocity) c->mrfFaceVel.assign(mrf_face_velocity, mrf_face_velocity + c->m.FT); else c->mrfFaceVel.clear();
    if (mrf_omega) c->mrfOmega.assign(mrf_omega, mrf_omega + 3 * (size_t)c->m.N); else c->mrfOmega.clear();
    return 0;
}

int orc_matrix_mul(Ctx* c, const double* xRho, const double* xRhoU, const double* xRhoE, double* yRho, double* yRhoU, double* yRhoE)
{
    if (!c->matrixSet) return fail(c, ICSB200_ESTATE, "matrix not assembled");
    const Mesh& m = c->m;
    size_t NT = (size_t)m.N + m.NB;
    vecd a(NT, 0.0), b(3 * NT, 0.0), e(NT, 0.0), ya, yb, ye;
    std::memcpy(a.data(), xRho, sizeof(double) * m.N);
    std::memcpy(b.data(), xRhoU, sizeof(double) * 3 * m.N);
    std::memcpy(e.data(), xRhoE, sizeof(double) * m.N);
    matrixMul(*c, a, b, e, ya, yb, ye);
    std::memcpy(yRho, ya.data(), sizeof(double) * m.N);
    std::memcpy(yRhoU, yb.data(), sizeof(double) * 3 * m.N);
    std::memcpy(yRhoE, ye.data(), sizeof(double) * m.N);
    return 0;
}

int orc_precondition(Ctx* c, int preconditioner, double* xRho, double* xRhoU, double* xRhoE)
{
    if (!c->matrixSet) return fail(c, ICSB200_ESTATE, "matrix not assembled");
    const int N = c->m.N;
    vecd a(xRho, xRho + N), b(xRhoU, xRhoU + 3 * (size_t)N), e(xRhoE, xRhoE + N);
    int r = precondition(*c, preconditioner, a, b, e);
    if (r) return fail(c, r, "All diagonals of coupledMatrix are zero.");
    std::memcpy(xRho, a.data(), sizeof(double) * N);
    std::memcpy(xRhoU, b.data(), sizeof(double) * 3 * N);
    std::memcpy(xRhoE, e.data(), sizeof(double) * N);
    return 0;
}

int orc_solve_delta(Ctx* c, const icsb200_solver_controls* ctl, double* dRho, double* dRhoU, double* dRhoE, icsb200_residuals* res)
{
    if (!c->matrixSet) return fail(c, ICSB200_ESTATE, "matrix not assembled");
    if (c->rhoPrev.empty()) { c->rhoPrev = c->rho; c->rhoUPrev = c->rhoU; c->rhoEPrev = c->rhoE; }
    int r = solveDelta(*c, *ctl, *res);
    if (r) return fail(c, r, "solveDelta failed");
    c->initRes = *res;
    c->haveInitRes = true;
    const int N = c->m.N;
    if (dRho) std::memcpy(dRho, c->dRho.data(), sizeof(double) * N);
    if (dRhoU) std::memcpy(dRhoU, c->dRhoU.data(), sizeof(double) * 3 * N);
    if (dRhoE) std::memcpy(dRhoE, c->dRhoE.data(), sizeof(double) * N);
    return 0;
}

int orc_update_fields(Ctx* c)
{
    boundLocalTimeStep(*c);
    updateFields(*c);
    c->firstIter = false;
    return 0;
}

int orc_iterate_dev(Ctx* c, const icsb200_solver_controls* ctl, icsb200_residuals* res)
{
    if (!c->stateSet) return fail(c, ICSB200_ESTATE, "state not set");
    int r = iterate(*c, *ctl, *res);
    if (r) return fail(c, r, "iterate failed");
    return 0;
}

// run n_iter outer iterations on every context of a world, one thread per context (the "MPI run")
int orc_world_iterate(Ctx** ctxs, int n, const icsb200_solver_controls* ctl, int n_iter, icsb200_residuals* res_out)
{
    std::vector<std::thread> th;
    std::vector<int> rc(n, 0);
    for (int r = 0; r < n; r++)
        th.emplace_back([&, r] {
            icsb200_residuals res;
            for (int it = 0; it < n_iter && rc[r] == 0; it++) rc[r] = iterate(*ctxs[r], *ctl, res);
            if (r == 0 && res_out) *res_out = res;
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < n; r++) if (rc[r]) return rc[r];
    return 0;
}

// SPMD setup helpers for worlds: mesh_set and state_set exchange halos, so they must run concurrently
int orc_world_mesh_set(Ctx** ctxs, int n, const int* n_cells, const int* n_internal_faces, const int* n_faces, const int* const* owner,
                       const int* const* neighbour, const double* const* Sf, const double* const* magSf, const double* const* weights,
                       const double* const* deltaCoeffs, const double* const* nonOrthDeltaCoeffs, const double* const* C, const double* const* V,
                       const double* const* Cf, const int* n_patches, const icsb200_patch* const* patches, const int solutionD[3])
{
    std::vector<std::thread> th;
    std::vector<int> rc(n, 0);
    for (int r = 0; r < n; r++)
        th.emplace_back([&, r] {
            rc[r] = orc_mesh_set(ctxs[r], n_cells[r], n_internal_faces[r], n_faces[r], owner[r], neighbour[r], Sf[r], magSf[r], weights[r],
                                 deltaCoeffs[r], nonOrthDeltaCoeffs[r], C[r], V[r], Cf[r], n_patches[r], patches[r], solutionD);
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < n; r++) if (rc[r]) return rc[r];
    return 0;
}

int orc_world_state_set(Ctx** ctxs, int n, const double* const* p, const double* const* U, const double* const* T)
{
    std::vector<std::thread> th;
    for (int r = 0; r < n; r++) th.emplace_back([&, r] { orc_state_set(ctxs[r], p[r], U[r], T[r]); });
    for (auto& t : th) t.join();
    return 0;
}

}  // extern "C"

// ---- test hooks (oracle only): expose the restated OpenFOAM operators for known-answer tests ----
extern "C" {

// fvc::interpolate(vf, pos_/neg_, limited scheme): cell values [N] (boundary = zeroGradient / coupled) -> L[FT], R[FT]
int orc_debug_reconstruct(Ctx* c, int lim, const double* cells, double* L, double* R)
{
    const Mesh& m = c->m;
    vecd vf((size_t)m.N + m.NB, 0.0), sl, sr;
    for (int i = 0; i < m.N; i++) vf[i] = cells[i];
    for (auto& p : m.patches)
        if (!m.empty(p)) for (int f = p.start; f < p.start + p.size; f++) vf[m.N + f - m.F] = cells[m.owner[f]];
    syncCoupled(*c, vf, 1);
    interpolateLimitedLR(*c, vf, lim, sl, sr);
    std::memcpy(L, sl.data(), sizeof(double) * m.FT);
    std::memcpy(R, sr.data(), sizeof(double) * m.FT);
    return 0;
}

// Gauss linear gradient of a cell field with zeroGradient boundaries -> grad[3N]
int orc_debug_grad(Ctx* c, const double* cells, double* grad)
{
    const Mesh& m = c->m;
    vecd vf((size_t)m.N + m.NB, 0.0), g;
    for (int i = 0; i < m.N; i++) vf[i] = cells[i];
    for (auto& p : m.patches)
        if (!m.empty(p)) for (int f = p.start; f < p.start + p.size; f++) vf[m.N + f - m.F] = cells[m.owner[f]];
    syncCoupled(*c, vf, 1);
    gradGauss(*c, vf, g);
    std::memcpy(grad, g.data(), sizeof(double) * 3 * m.N);
    return 0;
}

}  // extern "C"
