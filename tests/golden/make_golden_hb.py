"""Regenerates tests/golden/hb_box_roe_3instants.npz with the reference-structured HB oracle (oracle/oracle_hb.cpp).

Like the other fixtures it pins the oracle (and the CUDA path) against drift; it is not an output of the OpenFOAM binary
(parity unpinned).  Run from the repository root:  python -m tests.golden.make_golden_hb
"""
import os

import numpy as np

from icsfoam_b200 import cases
from oracle.pyoracle import HB

HERE = os.path.dirname(os.path.abspath(__file__))


def make_case():
    return cases.hb_box(4, 3, flux="ROE", seed=3)


def run(api_hb, case, NT):
    """api_hb: anything with the HB-flavoured calls (pyoracle.HB or tests.test_gpu_hb.GpuHB)."""
    out = {}
    api_hb.assemble()
    r = api_hb.residual()
    out.update(srcRho=r[0], srcRhoU=r[1], srcRhoE=r[2])
    out["rPseudoDeltaT"] = api_hb.pseudo()[0]
    for b in (0, 3, 8):
        out[f"diag{b}"] = api_hb.matrix_get_ldu(b)[0]
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(NT), rng.standard_normal((NT, 3)), rng.standard_normal(NT))
    y = api_hb.matrix_mul(*x)
    z = api_hb.precondition("LUSGS", *x)
    out.update(y0=y[0], y1=y[1], y2=y[2], z0=z[0], z1=z[1], z2=z[2])
    hist = []
    for _ in range(4):
        res = api_hb.iterate(case.controls)
        hist.append(list(res["s_init"]) + list(res["v_init"]) + [res["n_iterations"]])
    st = api_hb.state_get()
    out.update(history=np.array(hist), rho=st["rho"], rhoU=st["rhoU"], rhoE=st["rhoE"])
    return out


if __name__ == "__main__":
    case = make_case()
    data = run(HB(case), case, case.mesh.n_cells)
    np.savez_compressed(os.path.join(HERE, "hb_box_roe_3instants.npz"), **data)
    print("hb_box_roe_3instants", sum(v.nbytes for v in data.values()) // 1024, "KiB")
