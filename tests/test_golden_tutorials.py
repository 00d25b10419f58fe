"""Fixtures on the reference's own meshes (tests/golden/make_golden_tutorials.py): C2 forwardStep (polyhedral, 36 576 cells)
and C5 VKI-LS89 (28 059 cells, cyclic pair, viscous).  CPU: the oracle reproduces them; GPU: the CUDA path through the C ABI
reproduces them (checksums to 1e-12 for the reduction-free stages, solver stages to the north-star bars)."""
import os

import numpy as np
import pytest

from icsfoam_b200 import cases
from tests.common import GOLDEN
from tests.golden.make_golden_tutorials import TUTORIALS, run

EXACT = ["phi", "phiUp", "phiEp", "srcRho", "srcRhoU", "srcRhoE", "rPseudoDeltaT"]


def _compare(out, gold, tol_exact, tol_solve):
    for k in EXACT:
        ref = gold[k + "_chk"]
        # sum / abs sum / L2 / max: relative to the abs sum (a plain sum of signed fluxes cancels)
        scale = np.array([ref[1], ref[1], ref[2], ref[3]]) + 1e-300
        assert (np.abs(out[k + "_chk"] - ref) / scale).max() <= tol_exact, k
        s = gold[k + "_sample"]
        assert np.abs(out[k + "_sample"] - s).max() <= tol_exact * (np.abs(s).max() + 1e-300), k
    for k in ("dVByV_diag", "dVByV_upper", "dVByV_lower"):
        ref = gold[k + "_chk"]
        assert (np.abs(out[k + "_chk"] - ref) / (np.array([ref[1], ref[1], ref[2], ref[3]]) + 1e-300)).max() <= tol_exact, k
    assert out["restarts"][0] == gold["restarts"][0]
    assert np.array_equal(out["history"][:, -1], gold["history"][:, -1])
    assert np.allclose(out["history"][:, :5], gold["history"][:, :5], rtol=tol_solve, atol=1e-14)
    assert np.allclose(out["sInit"], gold["sInit"], rtol=tol_solve) and np.allclose(out["vInit"], gold["vInit"], rtol=tol_solve, atol=1e-14)
    for k in ("dRho", "dRhoU", "dRhoE", "rho", "rhoU", "rhoE"):
        ref = gold[k + "_chk"]
        assert (np.abs(out[k + "_chk"] - ref) / (np.array([ref[1], ref[1], ref[2], ref[3]]) + 1e-300)).max() <= tol_solve, k
    assert np.abs(out["rho_sample"] - gold["rho_sample"]).max() <= tol_solve * np.abs(gold["rho_sample"]).max()


def _case(name):
    tut, mk = TUTORIALS[name]
    d = cases.tutorial_dir(tut)
    if d is None:
        pytest.skip(f"{tut} tutorial not found ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
    return mk(d)


@pytest.mark.parametrize("name", sorted(TUTORIALS))
def test_oracle_matches_tutorial_fixture(name):
    from oracle.pyoracle import Oracle
    case = _case(name)
    o = case.apply(Oracle())
    _compare(run(o, case), np.load(os.path.join(GOLDEN, name + ".npz")), 1e-13, 1e-10)
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(TUTORIALS))
def test_gpu_matches_tutorial_fixture(gpu_context, name):
    case = _case(name)
    g = case.apply(gpu_context())
    _compare(run(g, case), np.load(os.path.join(GOLDEN, name + ".npz")), 1e-12, 1e-8)
