"""The outer-corrector bookkeeping of the standalone driver (icsfoam_b200/host/icsfoamB200.H `pseudotimeControl`) against a
transcription of the reference's pseudotimeControl::loop() / criteriaSatisfied() (src/cfdTools/pseudotimeControl/
pseudotimeControl.C:72-101,166-248, pseudotimeControlI.H:42-58) written here in Python: same sequence of loop() results and
"converged in N iterations" reports for steady and transient runs, including the extra final iteration after the criteria
are first met and the unchecked last allowed iteration."""
import os
import subprocess

import numpy as np
import pytest

from tests.conftest import ROOT

SRC = r"""
#include <iostream>
#include <sstream>
#include "icsfoamB200.H"
using namespace icsfoamB200;
int main(int argc, char** argv)
{
    const bool steady = std::atoi(argv[1]) != 0;
    const int nCorr = std::atoi(argv[2]), nCorrMin = std::atoi(argv[3]);
    const double tol = std::atof(argv[4]), tolRel = std::atof(argv[5]);
    const int nSteps = std::atoi(argv[6]);
    std::ostringstream log;
    pseudotimeControl ctl(steady, nCorr, nCorrMin, tol, tolRel, log);
    double res = 1.0;
    for (int step = 0; step < nSteps; step++) {
        bool notFinished;
        if (!steady) res = 1.0;
        while ((notFinished = ctl.loop())) {
            residualsIO r;
            for (double& v : r.sInitRes) v = res;
            for (double& v : r.vInitRes) v = 0.5 * res;
            ctl.setResidual(r);
            std::cout << "it " << step << " " << ctl.corr() << "\n";
            res *= 0.5;
            if (steady) break;
        }
        std::cout << "end " << step << " " << (notFinished ? 1 : 0) << "\n";
        if (steady && !notFinished) break;
    }
    std::cout << log.str();
    return 0;
}
"""


class RefControl:
    """pseudotimeControl.C transcribed: names as in the reference."""

    def __init__(self, steady, nCorrOuter, nCorrOuterMin, tol, tolRel):
        self.steadyState_, self.nCorrOuter_, self.nCorrOuterMin_, self.tol, self.tolRel = steady, nCorrOuter, nCorrOuterMin, tol, tolRel
        self.corr_, self.converged_, self.firstIteration_ = 0, False, True
        self.residuals_ = self.initResiduals_ = None
        self.log = []

    def finalIter(self):
        return self.converged_ or self.corr_ == self.nCorrOuter_

    def criteriaSatisfied(self):
        if self.firstIteration_ or self.corr_ == 1 or self.finalIter():
            self.firstIteration_ = False
            return False
        storeIni = self.corr_ == 2
        if storeIni:
            self.initResiduals_ = self.residuals_.copy()
        absCheck = self.residuals_.max() < self.tol
        relCheck = False
        if not storeIni:
            relCheck = (self.residuals_ / (self.initResiduals_ + 1e-150)).max() < self.tolRel
        return self.corr_ >= self.nCorrOuterMin_ and (absCheck or relCheck)

    def loop(self):
        self.corr_ += 1
        if not self.steadyState_ and self.corr_ == self.nCorrOuter_ + 1:
            if self.nCorrOuter_ != 1:
                self.log.append(f"pseudoTime: not converged within {self.nCorrOuter_} iterations")
            self.corr_ = 0
            return False
        completed = False
        if self.converged_ or self.criteriaSatisfied():
            if self.converged_:
                self.log.append(f"pseudoTime: converged in {self.corr_ - 1} iterations")
                self.corr_, self.converged_, completed = 0, False, True
            else:
                self.log.append(f"pseudoTime: iteration {self.corr_}")
                self.converged_ = True
        elif self.steadyState_ or self.corr_ <= self.nCorrOuter_:
            if self.nCorrOuter_ != 1:
                self.log.append(f"pseudoTime: iteration {self.corr_}")
        return not completed


def reference_trace(steady, nCorr, nCorrMin, tol, tolRel, nSteps):
    c = RefControl(steady, nCorr, nCorrMin, tol, tolRel)
    out, res = [], 1.0
    for step in range(nSteps):
        if not steady:
            res = 1.0
        while True:
            notFinished = c.loop()
            if not notFinished:
                break
            c.residuals_ = np.array([res, res, 0.5 * res, 0.5 * res, 0.5 * res])
            out.append(f"it {step} {c.corr_}")
            res *= 0.5
            if steady:
                break
        out.append(f"end {step} {1 if notFinished else 0}")
        if steady and not notFinished:
            break
    return out + c.log


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    d = tmp_path_factory.mktemp("ptc")
    src, exe = d / "ptc.cpp", d / "ptc"
    src.write_text(SRC)
    r = subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "icsfoam_b200", "host"), str(src), "-o", str(exe),
                        "-L", os.path.join(ROOT, "icsfoam_b200"), "-licsb200", "-Wl,-rpath," + os.path.join(ROOT, "icsfoam_b200"), "-Wl,-rpath,/usr/local/cuda/lib64", "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return str(exe)


@pytest.mark.parametrize("steady,nCorr,nCorrMin,tol,tolRel,nSteps", [
    (1, 20, 1, 1e-2, 1e-30, 40),     # steady, absolute criterion: converges, one final iteration, then "converged in"
    (1, 20, 1, 1e-30, 0.1, 40),      # steady, relative criterion (initial residuals of iteration 2)
    (1, 5, 1, 1e-30, 1e-30, 12),     # steady, never converges: nPseudoCorr is ignored, corr == nPseudoCorr is not checked
    (1, 20, 8, 0.3, 1e-30, 40),      # nPseudoCorrMin delays the check
    (0, 6, 1, 0.2, 1e-30, 3),        # transient, converges inside every time step
    (0, 4, 1, 1e-30, 1e-30, 3),      # transient, "not converged within"
    (0, 4, 1, 0.2, 1e-30, 2),        # transient, criteria met on the last allowed iteration: not checked there
    (0, 1, 1, 1e-30, 1e-30, 3),      # nPseudoCorr 1: silent single iteration per step
])
def test_loop_matches_the_reference_transcription(binary, steady, nCorr, nCorrMin, tol, tolRel, nSteps):
    r = subprocess.run([binary, str(steady), str(nCorr), str(nCorrMin), repr(tol), repr(tolRel), str(nSteps)], capture_output=True, text=True)
    assert r.returncode == 0
    got = [ln for ln in r.stdout.splitlines() if ln.strip()]
    want = reference_trace(bool(steady), nCorr, nCorrMin, tol, tolRel, nSteps)
    assert got == want
    if (steady, tol) == (1, 1e-2):   # 0.5^k < 1e-2 first at the residual of iteration 8 -> checked in loop 9, final iteration 9, report in loop 10
        assert "pseudoTime: converged in 9 iterations" in got and "it 8 9" in got and "it 9 10" not in got
