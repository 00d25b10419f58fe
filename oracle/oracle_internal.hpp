// oracle_internal.hpp — internal declarations of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle.hpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>

#include "oracle.hpp"

namespace orc {

// oracle_fields.cpp
void meshFinalize(Ctx& c);
void syncCoupled(Ctx& c, vecd& vf, int nc);
// phase-lag patches: overwrite the boundary slots of a LISTED field with the phase-lagged neighbour values
enum { LAG_P = 0, LAG_U, LAG_RHO, LAG_E, LAG_H, LAG_C0, LAG_C1, LAG_C2 /* c with derivedFields' cKind 0/1/2 */ };
void applyPhaseLag(Ctx& c, vecd& vf, int nc, int tag);
void resyncCoupledState(Ctx& c);
void interpolateLinear(const Ctx& c, const vecd& vf, vecd& sf);
void gradGauss(Ctx& c, const vecd& vf, vecd& grad);
void interpolateLimitedLR(Ctx& c, const vecd& vf, int lim, vecd& sfL, vecd& sfR);
void surfaceIntegrate(const Ctx& c, const vecd& ssf, int nc, vecd& div);
void correctBoundary(Ctx& c);
void valueInternalCoeffs(const Ctx& c, int pi, int f, double& pVIC, double uVIC[3], double& tVIC);
void gradientInternalCoeffs(const Ctx& c, int pi, int f, double uGIC[3], double& tGIC);
void stateInit(Ctx& c, const double* p, const double* U, const double* T);

// oracle_flux.cpp
void calcFlux(Ctx& c);
void residualsUpdate(Ctx& c);
void setCoAndDeltaT(Ctx& c);
void serUpdate(Ctx& c);
void pseudoDeltaT(Ctx& c);
void computeDdtCoeff(Ctx& c);
void createJacobian(Ctx& c);
void fullViscousJacobian(Ctx& c);
void updateFields(Ctx& c);
void boundLocalTimeStep(Ctx& c);
void newTimeStep(Ctx& c);

// oracle_solver.cpp
void Amul(const Ctx& c, const Blk& b, int rowDim, int colDim, const vecd& psi, vecd& Apsi);
int lusgsPrecondition(Ctx& c, const vecd& rD, vecd& sRho, vecd& vRhoU, vecd& sRhoE);
void luInverse(int n, const double* A, double* inv);
void matrixMul(Ctx& c, vecd& xRho, vecd& xRhoU, vecd& xRhoE, vecd& yRho, vecd& yRhoU, vecd& yRhoE);
int precondition(Ctx& c, int kind, vecd& xRho, vecd& xRhoU, vecd& xRhoE);
int solveDelta(Ctx& c, const icsb200_solver_controls& ctl, icsb200_residuals& res);
int iterate(Ctx& c, const icsb200_solver_controls& ctl, icsb200_residuals& res);

}  // namespace orc
