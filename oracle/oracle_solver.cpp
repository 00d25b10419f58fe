// oracle_solver.cpp — coupledMatrix product, LU-SGS / block-Jacobi preconditioners, restarted GMRES, outer iteration.
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
//
// Follows blockFvMatrix.C:329-601 (Amul), coupledMatrix.C:66-123 (matrixMul), lusgs.C:50-382,
// JacobiSmoother.C:42-203 + Jacobi.C:55-132, gmres.C:44-69 + 772-1110 (6-argument solveDelta),
// coupledMatrixSolver.C:198-221 (stop), residualsIO.H:204-240, coupledMatrix.C:297-385 (solveForIncr),
// outerLoop.H:51-99.
#include "oracle_internal.hpp"

namespace orc {

namespace {

// LduMatrix dot(): coefficient (nc doubles) times psi -> result; kind encoded by (rowDim, colDim)
// scalar.scalar -> scalar ; vector.vector -> scalar (&) ; vector.scalar -> vector (*) ; tensor.vector -> vector (&)
inline void dotAdd(const double* coef, int rowDim, int colDim, const double* x, double* y)
{
    if (rowDim == 1 && colDim == 1) y[0] += coef[0] * x[0];
    else if (rowDim == 1 && colDim == 3) y[0] += coef[0] * x[0] + coef[1] * x[1] + coef[2] * x[2];
    else if (rowDim == 3 && colDim == 1) { y[0] += coef[0] * x[0]; y[1] += coef[1] * x[0]; y[2] += coef[2] * x[0]; }
    else {
        y[0] += coef[0] * x[0] + coef[1] * x[1] + coef[2] * x[2];
        y[1] += coef[3] * x[0] + coef[4] * x[1] + coef[5] * x[2];
        y[2] += coef[6] * x[0] + coef[7] * x[1] + coef[8] * x[2];
    }
}
inline void dotSub(const double* coef, int rowDim, int colDim, const double* x, double* y)
{
    if (rowDim == 1 && colDim == 1) y[0] -= coef[0] * x[0];
    else if (rowDim == 1 && colDim == 3) y[0] -= coef[0] * x[0] + coef[1] * x[1] + coef[2] * x[2];
    else if (rowDim == 3 && colDim == 1) { y[0] -= coef[0] * x[0]; y[1] -= coef[1] * x[0]; y[2] -= coef[2] * x[0]; }
    else {
        y[0] -= coef[0] * x[0] + coef[1] * x[1] + coef[2] * x[2];
        y[1] -= coef[3] * x[0] + coef[4] * x[1] + coef[5] * x[2];
        y[2] -= coef[6] * x[0] + coef[7] * x[1] + coef[8] * x[2];
    }
}

}  // namespace

// blockFvMatrix::Amul — psi has boundary slots holding patchNeighbourField for coupled patches
void Amul(const Ctx& c, const Blk& b, int rowDim, int colDim, const vecd& psi, vecd& Apsi)
{
    const Mesh& m = c.m;
    const int nc = b.nc;
    Apsi.assign((size_t)rowDim * m.N, 0.0);
    for (int cell = 0; cell < m.N; cell++) dotAdd(&b.diag[(size_t)nc * cell], rowDim, colDim, &psi[(size_t)colDim * cell], &Apsi[(size_t)rowDim * cell]);
    if (b.hasOff)
        for (int face = 0; face < m.F; face++) {
            int u = m.neighbour[face], l = m.owner[face];
            dotAdd(&b.lower[(size_t)nc * face], rowDim, colDim, &psi[(size_t)colDim * l], &Apsi[(size_t)rowDim * u]);
            dotAdd(&b.upper[(size_t)nc * face], rowDim, colDim, &psi[(size_t)colDim * u], &Apsi[(size_t)rowDim * l]);
        }
    if (b.hasInt)
        for (auto& p : m.patches)
            if (m.coupled(p))
                for (int f = p.start; f < p.start + p.size; f++)
                    dotAdd(&b.intUpper[(size_t)nc * (f - m.F)], rowDim, colDim, &psi[(size_t)colDim * (m.N + f - m.F)], &Apsi[(size_t)rowDim * m.owner[f]]);
}

namespace {

struct BlkRef { int id, rowVar, colVar; };  // var: 0 rho, 1 rhoE (scalars), 2 rhoU (vector)
// coupledMatrix::matrixMul loop order: SS(i,j), SV(i,j), VS(i,j), VV
const BlkRef kOrder[9] = {{0, 0, 0}, {1, 0, 1}, {2, 1, 0}, {3, 1, 1}, {4, 0, 2}, {5, 1, 2}, {6, 2, 0}, {7, 2, 1}, {8, 2, 2}};

}  // namespace

// x vectors are [N+NB] / [3(N+NB)]; y vectors [N] / [3N]
void matrixMul(Ctx& c, vecd& xRho, vecd& xRhoU, vecd& xRhoE, vecd& yRho, vecd& yRhoU, vecd& yRhoE)
{
    const Mesh& m = c.m;
    syncCoupled(c, xRho, 1);
    syncCoupled(c, xRhoU, 3);
    syncCoupled(c, xRhoE, 1);
    yRho.assign(m.N, 0.0); yRhoU.assign(3 * (size_t)m.N, 0.0); yRhoE.assign(m.N, 0.0);
    vecd* xs[3] = {&xRho, &xRhoE, &xRhoU};
    vecd* ys[3] = {&yRho, &yRhoE, &yRhoU};
    vecd tmp;
    for (auto& r : kOrder) {
        const Blk& b = c.blk[r.id];
        if (!b.exists) continue;
        int rowDim = r.rowVar == 2 ? 3 : 1, colDim = r.colVar == 2 ? 3 : 1;
        Amul(c, b, rowDim, colDim, *xs[r.colVar], tmp);
        vecd& y = *ys[r.rowVar];
        for (size_t i = 0; i < tmp.size(); i++) y[i] += tmp[i];
    }
}

// ---------------------------------------------------------------------------------------------- LU-SGS
static int lusgsDiag(const Ctx& c, vecd& rD)
{
    const Mesh& m = c.m;
    rD.assign(m.N, GREAT);
    for (int celli = 0; celli < m.N; celli++) {
        for (int id : {0, 3}) {
            if (!c.blk[id].exists) return ICSB200_ESTATE;
            rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(c.blk[id].diag[celli]));
        }
        const double* dg = &c.blk[8].diag[9 * (size_t)celli];
        rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(dg[0]));
        rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(dg[4]));
        rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(dg[8]));
        if (rD[celli] < VSMALL) return ICSB200_ESINGULAR;
    }
    return 0;
}

int lusgsPrecondition(Ctx& c, const vecd& rD, vecd& sRho, vecd& vRhoU, vecd& sRhoE)
{
    const Mesh& m = c.m;
    double* vecs[3] = {sRho.data(), sRhoE.data(), vRhoU.data()};
    // lower sweep (lusgs.C:230-304)
    for (int celli = 0; celli < m.N; celli++) {
        double dStar[3][3];
        dStar[0][0] = rD[celli] * sRho[celli];
        dStar[1][0] = rD[celli] * sRhoE[celli];
        for (int d = 0; d < 3; d++) dStar[2][d] = rD[celli] * vRhoU[3 * (size_t)celli + d];
        for (auto& r : kOrder) {
            const Blk& b = c.blk[r.id];
            if (!b.exists || !b.hasOff) continue;
            int rowDim = r.rowVar == 2 ? 3 : 1, colDim = r.colVar == 2 ? 3 : 1;
            for (int k = m.cellFaceStart[celli]; k < m.cellFaceStart[celli + 1]; k++) {
                int faceI = m.cellFaces[k];
                if (faceI < m.F) {
                    int cellj = m.neighbour[faceI];
                    if (cellj > celli) dotSub(&b.lower[(size_t)b.nc * faceI], rowDim, colDim, dStar[r.colVar], vecs[r.rowVar] + (size_t)rowDim * cellj);
                }
            }
        }
    }
    // upper sweep (lusgs.C:306-381)
    for (int celli = m.N - 1; celli >= 0; celli--) {
        double dS[3][3];
        dS[0][0] = rD[celli] * sRho[celli];
        dS[1][0] = rD[celli] * sRhoE[celli];
        for (int d = 0; d < 3; d++) dS[2][d] = rD[celli] * vRhoU[3 * (size_t)celli + d];
        sRho[celli] = dS[0][0];
        sRhoE[celli] = dS[1][0];
        for (int d = 0; d < 3; d++) vRhoU[3 * (size_t)celli + d] = dS[2][d];
        for (auto& r : kOrder) {
            const Blk& b = c.blk[r.id];
            if (!b.exists || !b.hasOff) continue;
            int rowDim = r.rowVar == 2 ? 3 : 1, colDim = r.colVar == 2 ? 3 : 1;
            for (int k = m.cellFaceStart[celli]; k < m.cellFaceStart[celli + 1]; k++) {
                int faceI = m.cellFaces[k];
                if (faceI < m.F) {
                    int cellj = m.owner[faceI];
                    if (cellj < celli) dotSub(&b.upper[(size_t)b.nc * faceI], rowDim, colDim, dS[r.colVar], vecs[r.rowVar] + (size_t)rowDim * cellj);
                }
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------- Jacobi
// LUscalarMatrix::inv = LU decomposition with partial pivoting (Foam::LUDecompose) + column-by-column back substitution
void luInverse(int n, const double* A, double* inv)
{
    std::vector<vecd> a(n, vecd(n));
    std::vector<int> piv(n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a[i][j] = A[i * n + j];
    vecd vv(n);
    for (int i = 0; i < n; i++) {
        double largest = 0.0;
        for (int j = 0; j < n; j++) largest = std::max(largest, std::fabs(a[i][j]));
        vv[i] = 1.0 / largest;
    }
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < j; i++) { double sum = a[i][j]; for (int k = 0; k < i; k++) sum -= a[i][k] * a[k][j]; a[i][j] = sum; }
        int iMax = 0;
        double largest = 0.0;
        for (int i = j; i < n; i++) {
            double sum = a[i][j];
            for (int k = 0; k < j; k++) sum -= a[i][k] * a[k][j];
            a[i][j] = sum;
            double temp = vv[i] * std::fabs(sum);
            if (temp >= largest) { largest = temp; iMax = i; }
        }
        piv[j] = iMax;
        if (j != iMax) { for (int k = 0; k < n; k++) std::swap(a[iMax][k], a[j][k]); vv[iMax] = vv[j]; }
        if (a[j][j] == 0.0) a[j][j] = SMALL;
        if (j != n - 1) { double rDiag = 1.0 / a[j][j]; for (int i = j + 1; i < n; i++) a[i][j] *= rDiag; }
    }
    for (int col = 0; col < n; col++) {
        vecd x(n, 0.0);
        x[col] = 1.0;
        int ii = 0;
        for (int i = 0; i < n; i++) {
            int ip = piv[i];
            double sum = x[ip];
            x[ip] = x[i];
            if (ii != 0) for (int j = ii - 1; j < i; j++) sum -= a[i][j] * x[j];
            else if (sum != 0.0) ii = i + 1;
            x[i] = sum;
        }
        for (int i = n - 1; i >= 0; i--) {
            double sum = x[i];
            for (int j = i + 1; j < n; j++) sum -= a[i][j] * x[j];
            x[i] = sum / a[i][i];
        }
        for (int i = 0; i < n; i++) inv[i * n + col] = x[i];
    }
}

// Jacobi::precondition = one JacobiSmoother sweep from zero: x = D^-1 (b - (L+U) 0) = D^-1 b
// variable order of the dense block: scalars (rho, rhoE) then vector (rhoU)  (JacobiSmoother.C:51-91)
static int jacobiPrecondition(Ctx& c, vecd& sRho, vecd& vRhoU, vecd& sRhoE)
{
    const Mesh& m = c.m;
    for (int celli = 0; celli < m.N; celli++) {
        double J[5][5], inv[5][5];
        J[0][0] = c.blk[0].diag[celli]; J[0][1] = c.blk[1].diag[celli];
        J[1][0] = c.blk[2].diag[celli]; J[1][1] = c.blk[3].diag[celli];
        for (int d = 0; d < 3; d++) {
            J[0][2 + d] = c.blk[4].diag[3 * (size_t)celli + d];
            J[1][2 + d] = c.blk[5].diag[3 * (size_t)celli + d];
            J[2 + d][0] = c.blk[6].diag[3 * (size_t)celli + d];
            J[2 + d][1] = c.blk[7].diag[3 * (size_t)celli + d];
            for (int e = 0; e < 3; e++) J[2 + d][2 + e] = c.blk[8].diag[9 * (size_t)celli + 3 * d + e];
        }
        luInverse(5, &J[0][0], &inv[0][0]);
        // matrixMulNoDiag of a zero field is zero: sTmp = -(0 - source)
        double var[5] = {-(0.0 - sRho[celli]), -(0.0 - sRhoE[celli]), -(0.0 - vRhoU[3 * (size_t)celli]), -(0.0 - vRhoU[3 * (size_t)celli + 1]),
                         -(0.0 - vRhoU[3 * (size_t)celli + 2])};
        double res[5] = {0, 0, 0, 0, 0};
        for (int i = 0; i < 5; i++) for (int j = 0; j < 5; j++) res[i] += inv[i][j] * var[j];
        sRho[celli] = res[0]; sRhoE[celli] = res[1];
        for (int d = 0; d < 3; d++) vRhoU[3 * (size_t)celli + d] = res[2 + d];
    }
    return 0;
}

int precondition(Ctx& c, int kind, vecd& xRho, vecd& xRhoU, vecd& xRhoE)
{
    if (kind == ICSB200_PRECOND_LUSGS) {
        vecd rD;
        int e = lusgsDiag(c, rD);
        if (e) return e;
        return lusgsPrecondition(c, rD, xRho, xRhoU, xRhoE);
    }
    return jacobiPrecondition(c, xRho, xRhoU, xRhoE);
}

// ---------------------------------------------------------------------------------------------- GMRES
static inline void givensRotation(double h, double beta, double& cc, double& s)
{
    if (beta == 0) { cc = 1; s = 0; }
    else if (std::fabs(beta) > std::fabs(h)) { double tau = -h / beta; s = 1.0 / std::sqrt(1.0 + sqr(tau)); cc = s * tau; }
    else { double tau = -beta / h; cc = 1.0 / std::sqrt(1.0 + sqr(tau)); s = cc * tau; }
}

static double gSum(Ctx& c, const double* v, size_t n) { double s = 0; for (size_t i = 0; i < n; i++) s += v[i]; return c.comm->sum(s); }
static double gSumSqr(Ctx& c, const vecd& v, size_t n) { double s = 0; for (size_t i = 0; i < n; i++) s += v[i] * v[i]; return c.comm->sum(s); }
static double gSumMagSqrV(Ctx& c, const vecd& v, int N) { double s = 0; for (int i = 0; i < N; i++) s += v[3 * (size_t)i] * v[3 * (size_t)i] + v[3 * (size_t)i + 1] * v[3 * (size_t)i + 1] + v[3 * (size_t)i + 2] * v[3 * (size_t)i + 2]; return c.comm->sum(s); }
static double gSumProd(Ctx& c, const vecd& a, const vecd& b, size_t n) { double s = 0; for (size_t i = 0; i < n; i++) s += a[i] * b[i]; return c.comm->sum(s); }
static double gSumProdV(Ctx& c, const vecd& a, const vecd& b, int N) { double s = 0; for (int i = 0; i < N; i++) s += a[3 * (size_t)i] * b[3 * (size_t)i] + a[3 * (size_t)i + 1] * b[3 * (size_t)i + 1] + a[3 * (size_t)i + 2] * b[3 * (size_t)i + 2]; return c.comm->sum(s); }
static double gSumMag(Ctx& c, const vecd& v, size_t n) { double s = 0; for (size_t i = 0; i < n; i++) s += std::fabs(v[i]); return c.comm->sum(s); }

static double maxRes(const icsb200_residuals& r)
{
    double mv = -VGREAT;
    for (int i = 0; i < 2; i++) mv = std::max(mv, r.s_final[i]);
    mv = std::max(mv, std::max(r.v_final[0], std::max(r.v_final[1], r.v_final[2])));
    return mv;
}
static double maxRel(const icsb200_residuals& r, const int solD[3])
{
    double mv = -VGREAT;
    for (int i = 0; i < 2; i++) mv = std::max(mv, r.s_final[i] / (r.s_init[i] + ROOTVSMALL));
    for (int d = 0; d < 3; d++) if (solD[d] == 1) mv = std::max(mv, r.v_final[d] / (r.v_init[d] + ROOTVSMALL));
    return mv;
}
static bool stop(const Ctx& c, const icsb200_solver_controls& ctl, const icsb200_residuals& r)
{
    if (r.n_iterations < ctl.min_iter) return false;
    return (r.n_iterations >= ctl.max_iter) || (maxRes(r) < ctl.tolerance) || (maxRel(r, c.m.solutionD) < ctl.rel_tol);
}

// gmres::solveDelta, 6-argument form.  W = (rhoPrev, rhoUPrev, rhoEPrev); sources = c.src*; result in c.dRho/dRhoU/dRhoE
int solveDelta(Ctx& c, const icsb200_solver_controls& ctl, icsb200_residuals& res)
{
    const Mesh& m = c.m;
    const int N = m.N, nDirs = ctl.n_directions;
    const size_t NT = (size_t)N + m.NB;
    std::memset(&res, 0, sizeof(res));
    long long nTot = (long long)c.comm->sum((double)N);
    // dW = W (cells and boundary), minus its global average
    vecd dsRho(c.rhoPrev), dsRhoE(c.rhoEPrev), dvRhoU(c.rhoUPrev);
    dsRho.resize(NT); dsRhoE.resize(NT); dvRhoU.resize(3 * NT);
    {
        double avgRho = gSum(c, dsRho.data(), N) / nTot;
        for (size_t i = 0; i < NT; i++) dsRho[i] -= avgRho;
        double avgRhoE = gSum(c, dsRhoE.data(), N) / nTot;
        for (size_t i = 0; i < NT; i++) dsRhoE[i] -= avgRhoE;
        double sx = 0, sy = 0, sz = 0;
        for (int i = 0; i < N; i++) { sx += dvRhoU[3 * (size_t)i]; sy += dvRhoU[3 * (size_t)i + 1]; sz += dvRhoU[3 * (size_t)i + 2]; }
        double avg[3] = {c.comm->sum(sx) / nTot, c.comm->sum(sy) / nTot, c.comm->sum(sz) / nTot};
        for (size_t i = 0; i < NT; i++) for (int d = 0; d < 3; d++) dvRhoU[3 * i + d] -= avg[d];
    }
    vecd sTmp0, sTmp1, vTmp;
    matrixMul(c, dsRho, dvRhoU, dsRhoE, sTmp0, vTmp, sTmp1);
    std::fill(dsRho.begin(), dsRho.begin() + N, 0.0);
    std::fill(dsRhoE.begin(), dsRhoE.begin() + N, 0.0);
    std::fill(dvRhoU.begin(), dvRhoU.begin() + 3 * (size_t)N, 0.0);
    const vecd &sSrc0 = c.srcRho, &sSrc1 = c.srcRhoE, &vSrc = c.srcRhoU;
    double sNorm[2], vNorm;
    {
        double s = 0;
        for (int i = 0; i < N; i++) s += std::fabs(sTmp0[i]) + std::fabs(sSrc0[i]);
        sNorm[0] = c.comm->sum(s) + VSMALL;
        res.s_init[0] = gSumMag(c, sSrc0, N) / sNorm[0];
        s = 0;
        for (int i = 0; i < N; i++) s += std::fabs(sTmp1[i]) + std::fabs(sSrc1[i]);
        sNorm[1] = c.comm->sum(s) + VSMALL;
        res.s_init[1] = gSumMag(c, sSrc1, N) / sNorm[1];
        s = 0;
        auto mag3 = [](const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
        for (int i = 0; i < N; i++) s += mag3(&vTmp[3 * (size_t)i]) + mag3(&vSrc[3 * (size_t)i]);
        vNorm = c.comm->sum(s) + VSMALL;
        double cs[3] = {0, 0, 0};
        for (int i = 0; i < N; i++) for (int d = 0; d < 3; d++) cs[d] += std::fabs(vSrc[3 * (size_t)i + d]);
        for (int d = 0; d < 3; d++) res.v_init[d] = c.comm->sum(cs[d]) / vNorm;
        for (int i = 0; i < 2; i++) res.s_final[i] = res.s_init[i];
        for (int d = 0; d < 3; d++) res.v_final[d] = res.v_init[d];
    }
    // approximate initial residual: r0 = b
    sTmp0 = sSrc0; sTmp1 = sSrc1; vTmp = vSrc;
    std::vector<vecd> H(nDirs, vecd(nDirs, 0.0));
    vecd yh(nDirs, 0.0), bh(nDirs + 1, 0.0), cg(nDirs, 0.0), sg(nDirs, 0.0);
    std::vector<vecd> V0(nDirs, vecd(NT, 0.0)), V1(nDirs, vecd(NT, 0.0)), VV(nDirs, vecd(3 * NT, 0.0));
    vecd rD;
    if (ctl.preconditioner == ICSB200_PRECOND_LUSGS) { int e = lusgsDiag(c, rD); if (e) return e; }
    auto precon = [&](vecd& a, vecd& v, vecd& b) {
        if (ctl.preconditioner == ICSB200_PRECOND_LUSGS) lusgsPrecondition(c, rD, a, v, b);
        else jacobiPrecondition(c, a, v, b);
    };
    const bool smooth = ctl.solver == ICSB200_SOLVER_SMOOTH;
    do {
      if (smooth) {
        // smoothSolverCoupled::solveDelta (smoothSolverCoupled.C:449-515) with JacobiSmoother::smooth (JacobiSmoother.C:120-203):
        // nSweeps times  x <- D^-1 ( -(matrixMulNoDiag(x) - b) ).  matrixMulNoDiag is restated as matrixMul minus the product
        // with the diagonal blocks (variable order of the dense block: rho, rhoE, rhoU).
        for (int sweep = 0; sweep < nDirs; sweep++) {
            matrixMul(c, dsRho, dvRhoU, dsRhoE, sTmp0, vTmp, sTmp1);
            for (int k = 0; k < N; k++) {
                const size_t k3 = 3 * (size_t)k, k9 = 9 * (size_t)k;
                double d0 = c.blk[0].diag[k] * dsRho[k] + c.blk[1].diag[k] * dsRhoE[k];
                double d1 = c.blk[2].diag[k] * dsRho[k] + c.blk[3].diag[k] * dsRhoE[k];
                double dv[3];
                for (int d = 0; d < 3; d++) {
                    d0 += c.blk[4].diag[k3 + d] * dvRhoU[k3 + d];
                    d1 += c.blk[5].diag[k3 + d] * dvRhoU[k3 + d];
                    dv[d] = c.blk[6].diag[k3 + d] * dsRho[k] + c.blk[7].diag[k3 + d] * dsRhoE[k];
                    for (int e = 0; e < 3; e++) dv[d] += c.blk[8].diag[k9 + 3 * d + e] * dvRhoU[k3 + e];
                }
                sTmp0[k] = -((sTmp0[k] - d0) - sSrc0[k]);
                sTmp1[k] = -((sTmp1[k] - d1) - sSrc1[k]);
                for (int d = 0; d < 3; d++) vTmp[k3 + d] = -((vTmp[k3 + d] - dv[d]) - vSrc[k3 + d]);
            }
            sTmp0.resize(N); sTmp1.resize(N); vTmp.resize(3 * (size_t)N);
            jacobiPrecondition(c, sTmp0, vTmp, sTmp1);                 // D^-1 applied to the bracket (zero-start sweep == D^-1 b)
            for (int k = 0; k < N; k++) { dsRho[k] = sTmp0[k]; dsRhoE[k] = sTmp1[k]; }
            for (size_t k = 0; k < 3 * (size_t)N; k++) dvRhoU[k] = vTmp[k];
        }
      } else {
        precon(sTmp0, vTmp, sTmp1);
        double beta = 0.0;
        beta += gSumSqr(c, sTmp0, N);
        beta += gSumSqr(c, sTmp1, N);
        beta += gSumMagSqrV(c, vTmp, N);
        beta = std::sqrt(beta);
        std::fill(bh.begin(), bh.end(), 0.0);
        bh[0] = beta;
        for (int i = 0; i < nDirs; i++) {
            for (int k = 0; k < N; k++) { V0[i][k] = sTmp0[k] / beta; V1[i][k] = sTmp1[k] / beta; }
            for (size_t k = 0; k < 3 * (size_t)N; k++) VV[i][k] = vTmp[k] / beta;
            matrixMul(c, V0[i], VV[i], V1[i], sTmp0, vTmp, sTmp1);
            precon(sTmp0, vTmp, sTmp1);
            for (int j = 0; j <= i; j++) {
                beta = 0.0;
                beta += gSumProd(c, sTmp0, V0[j], N);
                beta += gSumProd(c, sTmp1, V1[j], N);
                beta += gSumProdV(c, vTmp, VV[j], N);
                H[j][i] = beta;
                for (int k = 0; k < N; k++) { sTmp0[k] -= H[j][i] * V0[j][k]; sTmp1[k] -= H[j][i] * V1[j][k]; }
                for (size_t k = 0; k < 3 * (size_t)N; k++) vTmp[k] -= H[j][i] * VV[j][k];
            }
            beta = 0.0;
            beta += gSumSqr(c, sTmp0, N);
            beta += gSumSqr(c, sTmp1, N);
            beta += gSumMagSqrV(c, vTmp, N);
            beta = std::sqrt(beta);
            for (int j = 0; j < i; j++) {
                const double Hji = H[j][i];
                H[j][i] = cg[j] * Hji - sg[j] * H[j + 1][i];
                H[j + 1][i] = sg[j] * Hji + cg[j] * H[j + 1][i];
            }
            givensRotation(H[i][i], beta, cg[i], sg[i]);
            const double bhi = bh[i];
            bh[i] = cg[i] * bhi - sg[i] * bh[i + 1];
            bh[i + 1] = sg[i] * bhi + cg[i] * bh[i + 1];
            H[i][i] = cg[i] * H[i][i] - sg[i] * beta;
        }
        for (int i = nDirs - 1; i >= 0; i--) {
            double sum = bh[i];
            for (int j = i + 1; j < nDirs; j++) sum -= H[i][j] * yh[j];
            yh[i] = sum / stabilise(H[i][i], VSMALL);
        }
        for (int i = 0; i < nDirs; i++) {
            const double yi = yh[i];
            for (int k = 0; k < N; k++) { dsRho[k] += yi * V0[i][k]; dsRhoE[k] += yi * V1[i][k]; }
            for (size_t k = 0; k < 3 * (size_t)N; k++) dvRhoU[k] += yi * VV[i][k];
        }
      }
        matrixMul(c, dsRho, dvRhoU, dsRhoE, sTmp0, vTmp, sTmp1);
        for (int k = 0; k < N; k++) { sTmp0[k] = sSrc0[k] - sTmp0[k]; sTmp1[k] = sSrc1[k] - sTmp1[k]; }
        for (size_t k = 0; k < 3 * (size_t)N; k++) vTmp[k] = vSrc[k] - vTmp[k];
        res.s_final[0] = gSumMag(c, sTmp0, N) / sNorm[0];
        res.s_final[1] = gSumMag(c, sTmp1, N) / sNorm[1];
        double cs[3] = {0, 0, 0};
        for (int k = 0; k < N; k++) for (int d = 0; d < 3; d++) cs[d] += std::fabs(vTmp[3 * (size_t)k + d]);
        for (int d = 0; d < 3; d++) { res.v_final[d] = c.comm->sum(cs[d]) / vNorm; if (m.solutionD[d] == -1) res.v_final[d] = 0.0; }
        res.n_iterations += smooth ? nDirs : 1;
    } while (!stop(c, ctl, res));
    // coupledMatrix::solveForIncr: zero the increment in non-solved directions (coupledMatrix.C:371-382)
    for (int d = 0; d < 3; d++)
        if (m.solutionD[d] == -1) for (int k = 0; k < N; k++) dvRhoU[3 * (size_t)k + d] = 0.0;
    c.dRho.assign(dsRho.begin(), dsRho.begin() + N);
    c.dRhoE.assign(dsRhoE.begin(), dsRhoE.begin() + N);
    c.dRhoU.assign(dvRhoU.begin(), dvRhoU.begin() + 3 * (size_t)N);
    return 0;
}

// one outer pseudo-time iteration: outerLoop.H:51-99 then updateFields.H (dbnsFoam.C:110-124)
int iterate(Ctx& c, const icsb200_solver_controls& ctl, icsb200_residuals& res)
{
    const Mesh& m = c.m;
    calcFlux(c);
    residualsUpdate(c);
    setCoAndDeltaT(c);
    computeDdtCoeff(c);
    c.rhoPrev.assign(c.rho.begin(), c.rho.end());
    c.rhoUPrev.assign(c.rhoU.begin(), c.rhoU.end());
    c.rhoEPrev.assign(c.rhoE.begin(), c.rhoE.end());
    createJacobian(c);
    int e = solveDelta(c, ctl, res);
    if (e) return e;
    c.initRes = res;
    c.haveInitRes = true;
    boundLocalTimeStep(c);
    updateFields(c);
    c.firstIter = false;
    (void)m;
    return 0;
}

}  // namespace orc
