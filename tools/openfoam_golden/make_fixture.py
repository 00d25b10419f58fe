#!/usr/bin/env python
"""Turn a time directory written by the reference (dbnsFoam + dumpGolden.H, see README.md here) into a committed fixture:

    python tools/openfoam_golden/make_fixture.py <tutorial> <refCase>/<time> <n_outer_iterations> tests/golden/openfoam_<tutorial>_<time>.npz

<tutorial> is one of shockTube / forwardStep / bump / VKI-LS89 (the case builders of icsfoam_b200/cases.py), <n_outer_iterations> the
number of outer pseudo-time iterations dbnsFoam had executed when it wrote the directory.  tests/test_openfoam_golden.py picks up every
tests/golden/openfoam_*.npz and holds the oracle to it (1e-12 on the face fluxes and matrix coefficients of the first iteration,
1e-8 on the fields) — which turns "parity unpinned" into "pinned" for that configuration."""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make(tutorial, time_dir, n_iter, out):
    fd = _load(os.path.join(HERE, "..", "foamdiff.py"), "foamdiff")
    cm = _load(os.path.join(HERE, "compare_matrix.py"), "compare_matrix")
    data = {"tutorial": np.array(tutorial), "n_iter": np.array(int(n_iter))}
    for name in ("p", "U", "T", "rho", "phi", "phiUp", "phiEp"):
        path = os.path.join(time_dir, name)
        if os.path.exists(path):
            arr, uniform = fd.read_field(path)["internal"]
            if not uniform:
                data[name] = arr[:, 0] if arr.shape[1] == 1 else arr
    eq = os.path.join(time_dir, "eqSystem")
    if os.path.isdir(eq):
        for blk in cm.BLOCKS:
            for part in ("diag", "upper", "lower"):
                path = os.path.join(eq, f"{blk}_{part}")
                if os.path.exists(path):
                    data[f"eq_{blk}_{part}"] = cm.read_list(path)
    np.savez_compressed(out, **data)
    return sorted(data)


if __name__ == "__main__":
    if len(sys.argv) != 5:
        sys.exit(__doc__)
    print("stored:", make(*sys.argv[1:]))
