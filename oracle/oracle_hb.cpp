// oracle_hb.cpp — Harmonic Balance: HBZone sources and blocks, the (2 nO, nO) coupled system of
// dbnsFullyImplicitHBFoam and its outer iteration.  TEST INFRASTRUCTURE ONLY (see oracle.hpp).  PARITY UNPINNED.
//
// Follows src/cfdTools/HB/HBZoneTemplates.C:38-92 (addSource), HBZone.C:435-518 (addBlock), HBZone.C:521-651 (cylindrical
// momentum source), applications/solvers/dbnsFullyImplicitHBFoam/{outerLoop.H, setCoAndDeltaT.H, residualsUpdate.H:72-74},
// and the generic (nScalar, nVector) loops of coupledMatrix.C:66-123, lusgs.C:50-382, JacobiSmoother.C:42-203,
// gmres.C:772-1110, coupledMatrixSolver.C:198-221.
//
// Structure kept from the reference: nO separate time-instance meshes/contexts ("subTimeLevelK"), one global system
// whose only inter-instance coupling is the diagonal V*D[J][K] of the (rho,rho), (rhoU,rhoU), (rhoE,rhoE) blocks.
// The CUDA product instead runs the nO instances as ONE mesh of nO disconnected copies; comparing the two is the test.
#include <thread>

#include "oracle_internal.hpp"

namespace orc {

struct HBZone {
    vecd D;        // nO x nO, row-major (HBZone::D_)
    bool allMesh;  // cellZoneID_ == -2
    veci cells;    // otherwise: the cellZone
    bool cyl;
    double axis[3], centre[3];
};

struct HBSys {
    int nO = 0;
    std::vector<Ctx*> inst;
    std::vector<HBZone> zones;
    std::vector<vecd> contSource, momSource, energySource;  // [nO][N], [nO][3N], [nO][N]
    std::vector<vecd> hbS;  // [J*nO+K][N]: diagonal of the off-instance blocks dSByS(2J,2K), dSByS(2J+1,2K+1); dVByV(J,K) = hbS * I
    bool haveInit = false, havePrev = false, firstIter = true, assembled = false;
    vecd sInit, vInit, sFinal, vFinal, sInitPrev, vInitPrev;  // [2 nO], [3 nO]
    int nIterations = 0;
    std::string err;
};

namespace {

inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

// HBZoneTemplates.C:38-92
void addSourceScalar(const HBSys& h, const HBZone& z, std::vector<vecd>& source, int varNum)
{
    const int nO = h.nO;
    const Mesh& m = h.inst[0]->m;
    auto var = [&](int K, int c) { return varNum == 0 ? h.inst[K]->rho[c] : h.inst[K]->rhoE[c]; };
    auto cell = [&](int celli) {
        for (int J = 0; J < nO; J++)
            for (int K = 0; K < nO; K++) source[J][celli] -= m.V[celli] * z.D[J * nO + K] * var(K, celli);
    };
    if (z.allMesh) {
        for (int J = 0; J < nO; J++) std::fill(source[J].begin(), source[J].end(), 0.0);
        for (int celli = 0; celli < m.N; celli++) cell(celli);
    } else
        for (int celli : z.cells) {
            for (int J = 0; J < nO; J++) source[J][celli] = 0.0;
            cell(celli);
        }
}

// HBZone.C:521-651 (cylindrical) / HBZoneTemplates.C:38-92 (Cartesian)
void addSourceVector(const HBSys& h, const HBZone& z, std::vector<vecd>& source)
{
    const int nO = h.nO;
    const Mesh& m = h.inst[0]->m;
    if (!z.cyl) {
        auto cell = [&](int celli) {
            for (int J = 0; J < nO; J++)
                for (int K = 0; K < nO; K++) {
                    const double vd = m.V[celli] * z.D[J * nO + K];
                    for (int d = 0; d < 3; d++) source[J][3 * (size_t)celli + d] -= vd * h.inst[K]->rhoU[3 * (size_t)celli + d];
                }
        };
        if (z.allMesh) {
            for (int J = 0; J < nO; J++) std::fill(source[J].begin(), source[J].end(), 0.0);
            for (int celli = 0; celli < m.N; celli++) cell(celli);
        } else
            for (int celli : z.cells) {
                for (int J = 0; J < nO; J++) for (int d = 0; d < 3; d++) source[J][3 * (size_t)celli + d] = 0.0;
                cell(celli);
            }
        return;
    }
    const double magAxis = std::sqrt(dot3(z.axis, z.axis));
    const double axisHat[3] = {z.axis[0] / magAxis, z.axis[1] / magAxis, z.axis[2] / magAxis};
    auto radial = [&](const Mesh& mk, int celli, double* rHat) {
        double r[3] = {mk.C[3 * (size_t)celli] - z.centre[0], mk.C[3 * (size_t)celli + 1] - z.centre[1], mk.C[3 * (size_t)celli + 2] - z.centre[2]};
        const double ar = dot3(axisHat, r);
        for (int d = 0; d < 3; d++) r[d] -= ar * axisHat[d];
        const double magr = std::sqrt(dot3(r, r));
        for (int d = 0; d < 3; d++) rHat[d] = r[d] / magr;
    };
    auto cell = [&](int celli) {
        for (int J = 0; J < nO; J++) {
            double sourceCyl[3] = {0, 0, 0};
            for (int K = 0; K < nO; K++) {
                double rHat[3], tHat[3];
                radial(h.inst[K]->m, celli, rHat);
                cross3(axisHat, rHat, tHat);
                const double* u = &h.inst[K]->rhoU[3 * (size_t)celli];
                const double UCyl[3] = {dot3(u, rHat), dot3(u, tHat), dot3(u, axisHat)};
                const double vd = m.V[celli] * z.D[J * nO + K];
                for (int d = 0; d < 3; d++) sourceCyl[d] += vd * UCyl[d];
            }
            double rHat[3], tHat[3];
            radial(h.inst[J]->m, celli, rHat);
            cross3(axisHat, rHat, tHat);
            for (int d = 0; d < 3; d++) {
                const double sourceCart = sourceCyl[0] * rHat[d] + sourceCyl[1] * tHat[d] + sourceCyl[2] * axisHat[d];
                source[J][3 * (size_t)celli + d] -= sourceCart;
            }
        }
    };
    if (z.allMesh) {
        for (int J = 0; J < nO; J++) std::fill(source[J].begin(), source[J].end(), 0.0);
        for (int celli = 0; celli < m.N; celli++) cell(celli);
    } else
        for (int celli : z.cells) {
            for (int J = 0; J < nO; J++) for (int d = 0; d < 3; d++) source[J][3 * (size_t)celli + d] = 0.0;
            cell(celli);
        }
}

// HB.addingSource x3 (outerLoop.H:28-30)
void addingSources(HBSys& h)
{
    for (auto& z : h.zones) addSourceScalar(h, z, h.contSource, 0);
    for (auto& z : h.zones) addSourceVector(h, z, h.momSource);
    for (auto& z : h.zones) addSourceScalar(h, z, h.energySource, 1);
}

// outerLoop.H:91-206 for every K (without the SER part): flux, residual + HB source, Jacobian, HB diagonal blocks
void assembleAll(HBSys& h)
{
    const int nO = h.nO;
    const int N = h.inst[0]->m.N;
    for (int K = 0; K < nO; K++) {
        Ctx& c = *h.inst[K];
        c.rhoPrev = c.rho; c.rhoUPrev = c.rhoU; c.rhoEPrev = c.rhoE;
    }
    for (int K = 0; K < nO; K++) {
        Ctx& c = *h.inst[K];
        const Mesh& m = c.m;
        calcFlux(c);
        residualsUpdate(c);  // dbnsFullyImplicitHBFoam/residualsUpdate.H:1-9 (+ fvm::ddt with the HBM inner scheme == 0)
        for (int i = 0; i < N; i++) {  // residualsUpdate.H:72-74
            c.srcRho[i] = c.srcRho[i] + h.contSource[K][i];
            for (int d = 0; d < 3; d++) c.srcRhoU[3 * (size_t)i + d] = c.srcRhoU[3 * (size_t)i + d] + h.momSource[K][3 * (size_t)i + d];
            c.srcRhoE[i] = c.srcRhoE[i] + h.energySource[K][i];
        }
        computeDdtCoeff(c);
        createJacobian(c);  // also copies the sources into the blocks
        // HB.addBlock(eqSystemBlock.dSByS(0,0) / dVByV(0,0) / dSByS(1,1), K, K)  (outerLoop.H:161-163, HBZone.C:435-518)
        for (auto& z : h.zones) {
            auto add = [&](int celli) {
                const double vd = m.V[celli] * z.D[K * nO + K];
                c.blk[0].diag[celli] += vd;
                c.blk[8].diag[9 * (size_t)celli + 0] += vd * 1.0;
                c.blk[8].diag[9 * (size_t)celli + 4] += vd * 1.0;
                c.blk[8].diag[9 * (size_t)celli + 8] += vd * 1.0;
                c.blk[3].diag[celli] += vd;
            };
            if (z.allMesh) for (int celli = 0; celli < N; celli++) add(celli);
            else for (int celli : z.cells) add(celli);
        }
        // off-instance blocks HBRho / HBRhoU / HBRhoE (outerLoop.H:178-204): diagonal only
        for (int J = 0; J < nO; J++) {
            if (J == K) continue;
            vecd& s = h.hbS[J * nO + K];
            s.assign(N, 0.0);
            for (auto& z : h.zones) {
                const Mesh& mj = h.inst[J]->m;
                if (z.allMesh) for (int celli = 0; celli < N; celli++) s[celli] += mj.V[celli] * z.D[J * nO + K];
                else for (int celli : z.cells) s[celli] += mj.V[celli] * z.D[J * nO + K];
            }
        }
    }
    h.assembled = true;
}

// coupledMatrix::matrixMul for the (2 nO, nO) system.  x: [nO][N+NB] / [nO][3(N+NB)], y: [nO][N] / [nO][3N].
// Blocks whose diagonal was set to Zero (outerLoop.H:193-203) contribute exact zeros and are skipped.
void hbMatrixMul(HBSys& h, std::vector<vecd>& xRho, std::vector<vecd>& xRhoU, std::vector<vecd>& xRhoE, std::vector<vecd>& yRho,
                 std::vector<vecd>& yRhoU, std::vector<vecd>& yRhoE)
{
    const int nO = h.nO;
    const int N = h.inst[0]->m.N;
    for (int I = 0; I < nO; I++) {
        syncCoupled(*h.inst[I], xRho[I], 1);
        syncCoupled(*h.inst[I], xRhoU[I], 3);
        syncCoupled(*h.inst[I], xRhoE[I], 1);
    }
    vecd tmp;
    for (int I = 0; I < nO; I++) {
        Ctx& c = *h.inst[I];
        for (int a = 0; a < 2; a++) {  // scalar row i = 2I + a
            vecd& y = a == 0 ? yRho[I] : yRhoE[I];
            y.assign(N, 0.0);
            for (int j = 0; j < 2 * nO; j++) {  // dSByS(i, j)
                const int J = j / 2, b = j % 2;
                const vecd& x = b == 0 ? xRho[J] : xRhoE[J];
                if (J == I) {
                    Amul(c, c.blk[a * 2 + b], 1, 1, x, tmp);
                    for (int k = 0; k < N; k++) y[k] += tmp[k];
                } else if (a == b) {
                    const vecd& s = h.hbS[I * nO + J];
                    for (int k = 0; k < N; k++) y[k] += s[k] * x[k];
                }
            }
            // dSByV(i, K): only K == I is populated
            Amul(c, c.blk[4 + a], 1, 3, xRhoU[I], tmp);
            for (int k = 0; k < N; k++) y[k] += tmp[k];
        }
    }
    for (int I = 0; I < nO; I++) {
        Ctx& c = *h.inst[I];
        vecd& y = yRhoU[I];
        y.assign(3 * (size_t)N, 0.0);
        for (int b = 0; b < 2; b++) {  // dVByS(I, 2I + b)
            Amul(c, c.blk[6 + b], 3, 1, b == 0 ? xRho[I] : xRhoE[I], tmp);
            for (size_t k = 0; k < 3 * (size_t)N; k++) y[k] += tmp[k];
        }
        for (int K = 0; K < nO; K++) {  // dVByV(I, K)
            if (K == I) {
                Amul(c, c.blk[8], 3, 3, xRhoU[I], tmp);
                for (size_t k = 0; k < 3 * (size_t)N; k++) y[k] += tmp[k];
            } else {
                const vecd& s = h.hbS[I * nO + K];
                for (int k = 0; k < N; k++)
                    for (int d = 0; d < 3; d++) y[3 * (size_t)k + d] += s[k] * xRhoU[K][3 * (size_t)k + d];
            }
        }
    }
}

// lusgs::lusgs (lusgs.C:50-123): ONE scalar per cell over the diagonals of all 2 nO scalar and nO vector variables
int hbLusgsDiag(const HBSys& h, vecd& rD)
{
    const int nO = h.nO;
    const int N = h.inst[0]->m.N;
    rD.assign(N, GREAT);
    for (int celli = 0; celli < N; celli++) {
        for (int i = 0; i < 2 * nO; i++) {
            const Blk& b = h.inst[i / 2]->blk[(i % 2) ? 3 : 0];
            rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(b.diag[celli]));
        }
        for (int I = 0; I < nO; I++) {
            const double* dg = &h.inst[I]->blk[8].diag[9 * (size_t)celli];
            rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(dg[0]));
            rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(dg[4]));
            rD[celli] = 1.0 / std::max(1.0 / rD[celli], std::fabs(dg[8]));
        }
        if (rD[celli] < VSMALL) return ICSB200_ESINGULAR;
    }
    return 0;
}

// lusgs::precondition (lusgs.C:220-382).  The off-instance blocks have neither lower nor upper (diagonal-only
// blockFvMatrix), so forwardSweep/reverseSweep skip them (lusgs.C:139,169): the sweep visits every cell once and, at
// that cell, updates each instance with that instance's own off-diagonals — the same operations, in the same order
// per target cell and variable, as one single-instance sweep per instance with the shared rDiagCoeff.
void hbLusgs(HBSys& h, const vecd& rD, std::vector<vecd>& sRho, std::vector<vecd>& vRhoU, std::vector<vecd>& sRhoE)
{
    for (int I = 0; I < h.nO; I++) lusgsPrecondition(*h.inst[I], rD, sRho[I], vRhoU[I], sRhoE[I]);
}

// Jacobi::precondition = one JacobiSmoother sweep from zero (JacobiSmoother.C:42-203) with the dense (2 nO + 3 nO)^2
// cell matrix.  DEVIATION: the reference writes the dVByV(v, nv) entries at column
// nScalar + nv (+1, +2), which overlaps for nVector > 1; the intended nScalar + 3 nv is used here.
void hbJacobi(HBSys& h, std::vector<vecd>& sRho, std::vector<vecd>& vRhoU, std::vector<vecd>& sRhoE)
{
    const int nO = h.nO, nS = 2 * nO, n = 5 * nO;
    const int N = h.inst[0]->m.N;
    vecd J((size_t)n * n), inv((size_t)n * n), var(n), res(n);
    for (int celli = 0; celli < N; celli++) {
        std::fill(J.begin(), J.end(), 0.0);
        for (int s = 0; s < nS; s++) {
            const int I = s / 2, a = s % 2;
            for (int ns = 0; ns < nS; ns++) {
                const int K = ns / 2, b = ns % 2;
                if (K == I) J[s * n + ns] = h.inst[I]->blk[a * 2 + b].diag[celli];
                else if (a == b) J[s * n + ns] = h.hbS[I * nO + K][celli];
            }
            for (int d = 0; d < 3; d++) J[s * n + nS + 3 * I + d] = h.inst[I]->blk[4 + a].diag[3 * (size_t)celli + d];
        }
        for (int v = 0; v < nO; v++) {
            for (int b = 0; b < 2; b++)
                for (int d = 0; d < 3; d++) J[(nS + 3 * v + d) * n + 2 * v + b] = h.inst[v]->blk[6 + b].diag[3 * (size_t)celli + d];
            for (int nv = 0; nv < nO; nv++)
                for (int d = 0; d < 3; d++)
                    for (int e = 0; e < 3; e++) {
                        double val;
                        if (nv == v) val = h.inst[v]->blk[8].diag[9 * (size_t)celli + 3 * d + e];
                        else val = d == e ? h.hbS[v * nO + nv][celli] : 0.0;
                        J[(nS + 3 * v + d) * n + nS + 3 * nv + e] = val;
                    }
        }
        luInverse(n, J.data(), inv.data());
        for (int I = 0; I < nO; I++) {
            var[2 * I] = -(0.0 - sRho[I][celli]);
            var[2 * I + 1] = -(0.0 - sRhoE[I][celli]);
            for (int d = 0; d < 3; d++) var[nS + 3 * I + d] = -(0.0 - vRhoU[I][3 * (size_t)celli + d]);
        }
        for (int i = 0; i < n; i++) {
            res[i] = 0.0;
            for (int j = 0; j < n; j++) res[i] += inv[(size_t)i * n + j] * var[j];
        }
        for (int I = 0; I < nO; I++) {
            sRho[I][celli] = res[2 * I];
            sRhoE[I][celli] = res[2 * I + 1];
            for (int d = 0; d < 3; d++) vRhoU[I][3 * (size_t)celli + d] = res[nS + 3 * I + d];
        }
    }
}

inline void givensRotation(double hh, double beta, double& cc, double& s)
{
    if (beta == 0) { cc = 1; s = 0; }
    else if (std::fabs(beta) > std::fabs(hh)) { double tau = -hh / beta; s = 1.0 / std::sqrt(1.0 + sqr(tau)); cc = s * tau; }
    else { double tau = -beta / hh; cc = 1.0 / std::sqrt(1.0 + sqr(tau)); s = cc * tau; }
}

// gmres::solveDelta (6-argument form, gmres.C:772-1110) with nScalar = 2 nO, nVector = nO
int hbSolveDelta(HBSys& h, const icsb200_solver_controls& ctl, std::vector<vecd>& dRho, std::vector<vecd>& dRhoU, std::vector<vecd>& dRhoE)
{
    const int nO = h.nO, nDirs = ctl.n_directions;
    Ctx& c0 = *h.inst[0];
    const Mesh& m = c0.m;
    const int N = m.N;
    const size_t NT = (size_t)N + m.NB;
    Comm* comm = c0.comm;
    const long long nTot = (long long)comm->sum((double)N);
    auto gsum = [&](double s) { return comm->sum(s); };
    // variable lists in the reference's order: scalars (rho_0, rhoE_0, rho_1, rhoE_1, ...) then vectors
    std::vector<vecd> dsRho(nO), dsRhoE(nO), dvRhoU(nO);
    for (int I = 0; I < nO; I++) {
        Ctx& c = *h.inst[I];
        dsRho[I] = c.rhoPrev; dsRhoE[I] = c.rhoEPrev; dvRhoU[I] = c.rhoUPrev;
        dsRho[I].resize(NT); dsRhoE[I].resize(NT); dvRhoU[I].resize(3 * NT);
    }
    for (int I = 0; I < nO; I++)
        for (int a = 0; a < 2; a++) {
            vecd& w = a == 0 ? dsRho[I] : dsRhoE[I];
            double s = 0;
            for (int i = 0; i < N; i++) s += w[i];
            const double avg = gsum(s) / nTot;
            for (size_t i = 0; i < NT; i++) w[i] -= avg;
        }
    for (int I = 0; I < nO; I++) {
        double sx = 0, sy = 0, sz = 0;
        vecd& w = dvRhoU[I];
        for (int i = 0; i < N; i++) { sx += w[3 * (size_t)i]; sy += w[3 * (size_t)i + 1]; sz += w[3 * (size_t)i + 2]; }
        const double avg[3] = {gsum(sx) / nTot, gsum(sy) / nTot, gsum(sz) / nTot};
        for (size_t i = 0; i < NT; i++) for (int d = 0; d < 3; d++) w[3 * i + d] -= avg[d];
    }
    std::vector<vecd> sTmp0(nO), sTmp1(nO), vTmp(nO);
    hbMatrixMul(h, dsRho, dvRhoU, dsRhoE, sTmp0, vTmp, sTmp1);
    for (int I = 0; I < nO; I++) {
        std::fill(dsRho[I].begin(), dsRho[I].begin() + N, 0.0);
        std::fill(dsRhoE[I].begin(), dsRhoE[I].begin() + N, 0.0);
        std::fill(dvRhoU[I].begin(), dvRhoU[I].begin() + 3 * (size_t)N, 0.0);
    }
    vecd sNorm(2 * nO), vNorm(nO);
    h.sInit.assign(2 * nO, 0.0); h.vInit.assign(3 * nO, 0.0); h.sFinal.assign(2 * nO, 0.0); h.vFinal.assign(3 * nO, 0.0);
    auto mag3 = [](const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    for (int I = 0; I < nO; I++) {
        const Ctx& c = *h.inst[I];
        for (int a = 0; a < 2; a++) {
            const vecd& t = a == 0 ? sTmp0[I] : sTmp1[I];
            const vecd& src = a == 0 ? c.srcRho : c.srcRhoE;
            double s = 0, sm = 0;
            for (int i = 0; i < N; i++) s += std::fabs(t[i]) + std::fabs(src[i]);
            sNorm[2 * I + a] = gsum(s) + VSMALL;
            for (int i = 0; i < N; i++) sm += std::fabs(src[i]);
            h.sInit[2 * I + a] = gsum(sm) / sNorm[2 * I + a];
            h.sFinal[2 * I + a] = h.sInit[2 * I + a];
        }
    }
    for (int I = 0; I < nO; I++) {
        const Ctx& c = *h.inst[I];
        double s = 0, cs[3] = {0, 0, 0};
        for (int i = 0; i < N; i++) s += mag3(&vTmp[I][3 * (size_t)i]) + mag3(&c.srcRhoU[3 * (size_t)i]);
        vNorm[I] = gsum(s) + VSMALL;
        for (int i = 0; i < N; i++) for (int d = 0; d < 3; d++) cs[d] += std::fabs(c.srcRhoU[3 * (size_t)i + d]);
        for (int d = 0; d < 3; d++) { h.vInit[3 * I + d] = gsum(cs[d]) / vNorm[I]; h.vFinal[3 * I + d] = h.vInit[3 * I + d]; }
    }
    for (int I = 0; I < nO; I++) { sTmp0[I] = h.inst[I]->srcRho; sTmp1[I] = h.inst[I]->srcRhoE; vTmp[I] = h.inst[I]->srcRhoU; }
    std::vector<vecd> H(nDirs, vecd(nDirs, 0.0));
    vecd yh(nDirs, 0.0), bh(nDirs + 1, 0.0), cg(nDirs, 0.0), sg(nDirs, 0.0);
    std::vector<std::vector<vecd>> V0(nDirs, std::vector<vecd>(nO, vecd(NT, 0.0))), V1(nDirs, std::vector<vecd>(nO, vecd(NT, 0.0))),
        VV(nDirs, std::vector<vecd>(nO, vecd(3 * NT, 0.0)));
    vecd rD;
    if (ctl.preconditioner == ICSB200_PRECOND_LUSGS) { int e = hbLusgsDiag(h, rD); if (e) return e; }
    auto precon = [&]() {
        if (ctl.preconditioner == ICSB200_PRECOND_LUSGS) hbLusgs(h, rD, sTmp0, vTmp, sTmp1);
        else hbJacobi(h, sTmp0, vTmp, sTmp1);
    };
    auto sumSqrAll = [&]() {
        double beta = 0.0;
        for (int I = 0; I < nO; I++)
            for (int a = 0; a < 2; a++) {
                const vecd& t = a == 0 ? sTmp0[I] : sTmp1[I];
                double s = 0;
                for (int k = 0; k < N; k++) s += t[k] * t[k];
                beta += gsum(s);
            }
        for (int I = 0; I < nO; I++) {
            const vecd& t = vTmp[I];
            double s = 0;
            for (int k = 0; k < N; k++) s += t[3 * (size_t)k] * t[3 * (size_t)k] + t[3 * (size_t)k + 1] * t[3 * (size_t)k + 1] + t[3 * (size_t)k + 2] * t[3 * (size_t)k + 2];
            beta += gsum(s);
        }
        return beta;
    };
    h.nIterations = 0;
    bool stopNow;
    do {
        precon();
        double beta = std::sqrt(sumSqrAll());
        std::fill(bh.begin(), bh.end(), 0.0);
        bh[0] = beta;
        for (int i = 0; i < nDirs; i++) {
            for (int I = 0; I < nO; I++) {
                for (int k = 0; k < N; k++) { V0[i][I][k] = sTmp0[I][k] / beta; V1[i][I][k] = sTmp1[I][k] / beta; }
                for (size_t k = 0; k < 3 * (size_t)N; k++) VV[i][I][k] = vTmp[I][k] / beta;
            }
            hbMatrixMul(h, V0[i], VV[i], V1[i], sTmp0, vTmp, sTmp1);
            precon();
            for (int j = 0; j <= i; j++) {
                beta = 0.0;
                for (int I = 0; I < nO; I++)
                    for (int a = 0; a < 2; a++) {
                        const vecd& t = a == 0 ? sTmp0[I] : sTmp1[I];
                        const vecd& v = a == 0 ? V0[j][I] : V1[j][I];
                        double s = 0;
                        for (int k = 0; k < N; k++) s += t[k] * v[k];
                        beta += gsum(s);
                    }
                for (int I = 0; I < nO; I++) {
                    const vecd &t = vTmp[I], &v = VV[j][I];
                    double s = 0;
                    for (int k = 0; k < N; k++) s += t[3 * (size_t)k] * v[3 * (size_t)k] + t[3 * (size_t)k + 1] * v[3 * (size_t)k + 1] + t[3 * (size_t)k + 2] * v[3 * (size_t)k + 2];
                    beta += gsum(s);
                }
                H[j][i] = beta;
                for (int I = 0; I < nO; I++) {
                    for (int k = 0; k < N; k++) { sTmp0[I][k] -= H[j][i] * V0[j][I][k]; sTmp1[I][k] -= H[j][i] * V1[j][I][k]; }
                    for (size_t k = 0; k < 3 * (size_t)N; k++) vTmp[I][k] -= H[j][i] * VV[j][I][k];
                }
            }
            beta = std::sqrt(sumSqrAll());
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
            for (int I = 0; I < nO; I++) {
                for (int k = 0; k < N; k++) { dsRho[I][k] += yi * V0[i][I][k]; dsRhoE[I][k] += yi * V1[i][I][k]; }
                for (size_t k = 0; k < 3 * (size_t)N; k++) dvRhoU[I][k] += yi * VV[i][I][k];
            }
        }
        hbMatrixMul(h, dsRho, dvRhoU, dsRhoE, sTmp0, vTmp, sTmp1);
        for (int I = 0; I < nO; I++) {
            const Ctx& c = *h.inst[I];
            for (int k = 0; k < N; k++) { sTmp0[I][k] = c.srcRho[k] - sTmp0[I][k]; sTmp1[I][k] = c.srcRhoE[k] - sTmp1[I][k]; }
            for (size_t k = 0; k < 3 * (size_t)N; k++) vTmp[I][k] = c.srcRhoU[k] - vTmp[I][k];
        }
        for (int I = 0; I < nO; I++)
            for (int a = 0; a < 2; a++) {
                const vecd& t = a == 0 ? sTmp0[I] : sTmp1[I];
                double s = 0;
                for (int k = 0; k < N; k++) s += std::fabs(t[k]);
                h.sFinal[2 * I + a] = gsum(s) / sNorm[2 * I + a];
            }
        for (int I = 0; I < nO; I++) {
            double cs[3] = {0, 0, 0};
            for (int k = 0; k < N; k++) for (int d = 0; d < 3; d++) cs[d] += std::fabs(vTmp[I][3 * (size_t)k + d]);
            for (int d = 0; d < 3; d++) {
                h.vFinal[3 * I + d] = gsum(cs[d]) / vNorm[I];
                if (m.solutionD[d] == -1) h.vFinal[3 * I + d] = 0.0;
            }
        }
        h.nIterations++;
        // solver::stop (coupledMatrixSolver.C:198-221) with residualsIO::max / maxRel over all variables
        if (h.nIterations < ctl.min_iter) stopNow = false;
        else {
            double mx = -VGREAT, mr = -VGREAT;
            for (int i = 0; i < 2 * nO; i++) { mx = std::max(mx, h.sFinal[i]); mr = std::max(mr, h.sFinal[i] / (h.sInit[i] + ROOTVSMALL)); }
            for (int I = 0; I < nO; I++) {
                mx = std::max(mx, std::max(h.vFinal[3 * I], std::max(h.vFinal[3 * I + 1], h.vFinal[3 * I + 2])));
                for (int d = 0; d < 3; d++) if (m.solutionD[d] == 1) mr = std::max(mr, h.vFinal[3 * I + d] / (h.vInit[3 * I + d] + ROOTVSMALL));
            }
            stopNow = (h.nIterations >= ctl.max_iter) || (mx < ctl.tolerance) || (mr < ctl.rel_tol);
        }
    } while (!stopNow);
    dRho.resize(nO); dRhoU.resize(nO); dRhoE.resize(nO);
    for (int I = 0; I < nO; I++) {
        for (int d = 0; d < 3; d++)
            if (m.solutionD[d] == -1) for (int k = 0; k < N; k++) dvRhoU[I][3 * (size_t)k + d] = 0.0;
        dRho[I].assign(dsRho[I].begin(), dsRho[I].begin() + N);
        dRhoE[I].assign(dsRhoE[I].begin(), dsRhoE[I].begin() + N);
        dRhoU[I].assign(dvRhoU[I].begin(), dvRhoU[I].begin() + 3 * (size_t)N);
    }
    return 0;
}

// one outer iteration of dbnsFullyImplicitHBFoam (outerLoop.H)
int hbIterate(HBSys& h, const icsb200_solver_controls& ctl)
{
    const int nO = h.nO;
    const icsb200_schemes& sch = h.inst[0]->sch;
    addingSources(h);
    // SER (outerLoop.H:32-64).  As coded, coNumRatio stays 0 on the first iteration that has an initRes but no
    // prevRes yet, and setCoAndDeltaT.H:3-11 then multiplies the Courant field by it (-> clamped to pseudoCoNumMin).
    double coNumRatio = 0;
    bool applyRatio = false;
    if (h.haveInit) {
        if (!h.firstIter && h.havePrev) {
            double normInitSqr = 0, normPrevSqr = 0;
            for (int J = 0; J < nO; J++) {
                normInitSqr += sqr(h.sInit[2 * J]) + sqr(h.sInit[2 * J + 1]) + (h.vInit[3 * J] * h.vInit[3 * J] + h.vInit[3 * J + 1] * h.vInit[3 * J + 1] + h.vInit[3 * J + 2] * h.vInit[3 * J + 2]);
                normPrevSqr += sqr(h.sInitPrev[2 * J]) + sqr(h.sInitPrev[2 * J + 1]) + (h.vInitPrev[3 * J] * h.vInitPrev[3 * J] + h.vInitPrev[3 * J + 1] * h.vInitPrev[3 * J + 1] + h.vInitPrev[3 * J + 2] * h.vInitPrev[3 * J + 2]);
            }
            const double normInit = std::sqrt(normInitSqr), normPrev = std::sqrt(normPrevSqr);
            coNumRatio = normPrev / normInit;
            coNumRatio = std::max(std::min(coNumRatio, sch.pseudo_co_num_max_incr), sch.pseudo_co_num_min_decr);
        }
        h.sInitPrev = h.sInit; h.vInitPrev = h.vInit;
        h.havePrev = true;
        applyRatio = !h.firstIter && h.havePrev;  // setCoAndDeltaT.H:3-5 (prevRes has just been set)
    }
    for (int K = 0; K < nO; K++) {
        Ctx& c = *h.inst[K];
        if (applyRatio)
            for (auto& v : c.pseudoCoField) { v *= coNumRatio; v = std::max(std::min(v, sch.pseudo_co_num_max), sch.pseudo_co_num_min); }
        pseudoDeltaT(c);  // always local time stepping in the HB solver
    }
    assembleAll(h);
    std::vector<vecd> dRho, dRhoU, dRhoE;
    int e = hbSolveDelta(h, ctl, dRho, dRhoU, dRhoE);
    if (e) return e;
    h.haveInit = true;
    for (int K = 0; K < nO; K++) {
        Ctx& c = *h.inst[K];
        c.dRho = dRho[K]; c.dRhoU = dRhoU[K]; c.dRhoE = dRhoE[K];
        boundLocalTimeStep(c);
        updateFields(c);
        c.firstIter = false;
    }
    // phase-lag patches read every instance: refresh the patchNeighbourField caches once all instances are updated
    for (int K = 0; K < nO; K++) if (!h.inst[K]->lagRow.empty()) resyncCoupledState(*h.inst[K]);
    h.firstIter = false;
    return 0;
}

}  // namespace

}  // namespace orc

using namespace orc;

extern "C" {

void* orc_hb_create(Ctx** inst, int nO)
{
    HBSys* h = new HBSys;
    h->nO = nO;
    h->inst.assign(inst, inst + nO);
    const int N = inst[0]->m.N;
    h->contSource.assign(nO, vecd(N, 0.0));
    h->momSource.assign(nO, vecd(3 * (size_t)N, 0.0));
    h->energySource.assign(nO, vecd(N, 0.0));
    h->hbS.assign((size_t)nO * nO, vecd(N, 0.0));
    return h;
}
void orc_hb_destroy(void* hb) { delete (HBSys*)hb; }

// n_cells == -2: the zone covers the whole mesh (allMesh); otherwise cells[] is the cellZone
int orc_hb_zone_add(void* hb, const double* D, int n_cells, const int* cells, int cyl, const double* axis, const double* centre)
{
    HBSys& h = *(HBSys*)hb;
    HBZone z;
    z.D.assign(D, D + (size_t)h.nO * h.nO);
    z.allMesh = n_cells == -2;
    if (!z.allMesh) z.cells.assign(cells, cells + n_cells);
    z.cyl = cyl != 0;
    for (int d = 0; d < 3; d++) { z.axis[d] = axis ? axis[d] : (d == 2); z.centre[d] = centre ? centre[d] : 0.0; }
    h.zones.push_back(z);
    return 0;
}

// arrays are instance-major: [nO][N], [nO][3N]
int orc_hb_sources(void* hb, double* cont, double* mom, double* energy)
{
    HBSys& h = *(HBSys*)hb;
    const int N = h.inst[0]->m.N;
    addingSources(h);
    for (int J = 0; J < h.nO; J++) {
        if (cont) std::memcpy(cont + (size_t)J * N, h.contSource[J].data(), sizeof(double) * N);
        if (mom) std::memcpy(mom + (size_t)J * 3 * N, h.momSource[J].data(), sizeof(double) * 3 * N);
        if (energy) std::memcpy(energy + (size_t)J * N, h.energySource[J].data(), sizeof(double) * N);
    }
    return 0;
}

// sources + pseudo time step (no SER) + flux, residual, Jacobian and HB blocks of every instance
int orc_hb_assemble(void* hb)
{
    HBSys& h = *(HBSys*)hb;
    addingSources(h);
    for (Ctx* c : h.inst) pseudoDeltaT(*c);
    assembleAll(h);
    return 0;
}

static void splitIn(const HBSys& h, const double* a, int nc, std::vector<vecd>& out, bool withBoundary)
{
    const Mesh& m = h.inst[0]->m;
    const size_t n = (size_t)nc * m.N, nt = (size_t)nc * ((size_t)m.N + (withBoundary ? m.NB : 0));
    out.assign(h.nO, vecd(nt, 0.0));
    for (int I = 0; I < h.nO; I++) std::memcpy(out[I].data(), a + I * n, sizeof(double) * n);
}
static void joinOut(const HBSys& h, const std::vector<vecd>& in, int nc, double* a)
{
    const size_t n = (size_t)nc * h.inst[0]->m.N;
    for (int I = 0; I < h.nO; I++) std::memcpy(a + I * n, in[I].data(), sizeof(double) * n);
}

int orc_hb_matrix_mul(void* hb, const double* xRho, const double* xRhoU, const double* xRhoE, double* yRho, double* yRhoU, double* yRhoE)
{
    HBSys& h = *(HBSys*)hb;
    if (!h.assembled) return ICSB200_ESTATE;
    std::vector<vecd> a, b, e, ya(h.nO), yb(h.nO), ye(h.nO);
    splitIn(h, xRho, 1, a, true); splitIn(h, xRhoU, 3, b, true); splitIn(h, xRhoE, 1, e, true);
    hbMatrixMul(h, a, b, e, ya, yb, ye);
    joinOut(h, ya, 1, yRho); joinOut(h, yb, 3, yRhoU); joinOut(h, ye, 1, yRhoE);
    return 0;
}

int orc_hb_precondition(void* hb, int preconditioner, double* xRho, double* xRhoU, double* xRhoE)
{
    HBSys& h = *(HBSys*)hb;
    if (!h.assembled) return ICSB200_ESTATE;
    std::vector<vecd> a, b, e;
    splitIn(h, xRho, 1, a, false); splitIn(h, xRhoU, 3, b, false); splitIn(h, xRhoE, 1, e, false);
    if (preconditioner == ICSB200_PRECOND_LUSGS) {
        vecd rD;
        int r = hbLusgsDiag(h, rD);
        if (r) return r;
        hbLusgs(h, rD, a, b, e);
    } else hbJacobi(h, a, b, e);
    joinOut(h, a, 1, xRho); joinOut(h, b, 3, xRhoU); joinOut(h, e, 1, xRhoE);
    return 0;
}

int orc_hb_solve_delta(void* hb, const icsb200_solver_controls* ctl, double* dRho, double* dRhoU, double* dRhoE)
{
    HBSys& h = *(HBSys*)hb;
    if (!h.assembled) return ICSB200_ESTATE;
    std::vector<vecd> a, b, e;
    int r = hbSolveDelta(h, *ctl, a, b, e);
    if (r) return r;
    h.haveInit = true;
    if (dRho) joinOut(h, a, 1, dRho);
    if (dRhoU) joinOut(h, b, 3, dRhoU);
    if (dRhoE) joinOut(h, e, 1, dRhoE);
    return 0;
}

int orc_hb_iterate(void* hb, const icsb200_solver_controls* ctl, int n_iter)
{
    HBSys& h = *(HBSys*)hb;
    for (int it = 0; it < n_iter; it++) {
        int r = hbIterate(h, *ctl);
        if (r) return r;
    }
    return 0;
}

// the sources of the global system after orc_hb_assemble / inside an iteration: R*V + HB source, instance-major
// P ranks, each with its own HB system over its partition (instance contexts attached to the same World): one thread per
// rank runs n_iter outer iterations SPMD — halos per instance through the mailboxes, reductions in rank order
int orc_world_hb_iterate(void** hbs, int n, const icsb200_solver_controls* ctl, int n_iter)
{
    std::vector<std::thread> th;
    std::vector<int> rc(n, 0);
    for (int r = 0; r < n; r++)
        th.emplace_back([&, r] {
            HBSys& h = *(HBSys*)hbs[r];
            for (int it = 0; it < n_iter && rc[r] == 0; it++) rc[r] = hbIterate(h, *ctl);
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < n; r++) if (rc[r]) return rc[r];
    return 0;
}

// phase-lag cyclic patch `patch` (index in the instance mesh): Dpl = Re(EInv M(IBPA) E), n x n row-major, for THIS side of the pair
// (owner +IBPA, neighbour -IBPA; phaseLagCyclicFvPatchField.C:166-330).  Call once per patch after all instances are initialised.
int orc_hb_phaselag_set(void* hb, int patch, const double* Dpl)
{
    HBSys& h = *(HBSys*)hb;
    const int nO = h.nO;
    for (int K = 0; K < nO; K++) {
        Ctx& c = *h.inst[K];
        c.hbSiblings.assign(h.inst.begin(), h.inst.end());
        c.hbIndex = K;
        c.lagRow[patch] = vecd(Dpl + (size_t)K * nO, Dpl + (size_t)(K + 1) * nO);
    }
    for (int K = 0; K < nO; K++) resyncCoupledState(*h.inst[K]);
    return 0;
}

int orc_hb_system_sources(void* hb, double* sRho, double* sRhoU, double* sRhoE)
{
    HBSys& h = *(HBSys*)hb;
    const int N = h.inst[0]->m.N;
    for (int J = 0; J < h.nO; J++) {
        const Ctx& c = *h.inst[J];
        if (sRho) std::memcpy(sRho + (size_t)J * N, c.srcRho.data(), sizeof(double) * N);
        if (sRhoU) std::memcpy(sRhoU + (size_t)J * 3 * N, c.srcRhoU.data(), sizeof(double) * 3 * N);
        if (sRhoE) std::memcpy(sRhoE + (size_t)J * N, c.srcRhoE.data(), sizeof(double) * N);
    }
    return 0;
}

// pseudo time-step fields of all instances, instance-major
int orc_hb_pseudo(void* hb, double* rPseudoDeltaT, double* pseudoCo)
{
    HBSys& h = *(HBSys*)hb;
    const int N = h.inst[0]->m.N;
    for (int J = 0; J < h.nO; J++) {
        const Ctx& c = *h.inst[J];
        if (rPseudoDeltaT) std::memcpy(rPseudoDeltaT + (size_t)J * N, c.rPseudoDeltaT.data(), sizeof(double) * N);
        if (pseudoCo) std::memcpy(pseudoCo + (size_t)J * N, c.pseudoCoField.data(), sizeof(double) * N);
    }
    return 0;
}

// residualsIO of the last solve: scalars [2 nO] (rho_0, rhoE_0, rho_1, ...), vectors [3 nO]
int orc_hb_residuals_get(void* hb, double* sInit, double* vInit, double* sFinal, double* vFinal, int* nIterations)
{
    HBSys& h = *(HBSys*)hb;
    if (h.sInit.empty()) return ICSB200_ESTATE;
    std::memcpy(sInit, h.sInit.data(), sizeof(double) * 2 * h.nO);
    std::memcpy(vInit, h.vInit.data(), sizeof(double) * 3 * h.nO);
    std::memcpy(sFinal, h.sFinal.data(), sizeof(double) * 2 * h.nO);
    std::memcpy(vFinal, h.vFinal.data(), sizeof(double) * 3 * h.nO);
    if (nIterations) *nIterations = h.nIterations;
    return 0;
}

}  // extern "C"
