#!/usr/bin/env python
"""Compare the LDU arrays of the coupledMatrix sub-blocks dumped by the reference (dumpGolden.H) and by `dbnsB200 -writeFlux`:

    python tools/openfoam_golden/compare_matrix.py <referenceCase>/<time>/eqSystem <b200Case>/<time>/eqSystem [--rtol 1e-12]

Every file is an OpenFOAM ASCII list (`N ( ... )`, or `N{value}` for a uniform list) of scalars, vectors or tensors.  The error of
a block is taken relative to the largest coefficient of that block (diag, upper and lower together)."""
import argparse
import os
import re
import sys

import numpy as np

BLOCKS = ["dSByS_0_0", "dSByS_0_1", "dSByS_1_0", "dSByS_1_1", "dSByV_0_0", "dSByV_1_0", "dVByS_0_0", "dVByS_0_1", "dVByV_0_0"]


def read_list(path):
    txt = open(path).read()
    h = txt.find("FoamFile")
    if h >= 0:
        txt = txt[txt.index("}", h) + 1:]
    m = re.match(r"\s*(\d+)\s*\{([^}]*)\}", txt)              # uniform list N{v}
    if m:
        v = np.array(m.group(2).replace("(", " ").replace(")", " ").split(), float)
        return np.tile(v, (int(m.group(1)), 1)).reshape(-1)
    m = re.match(r"\s*(\d+)\s*\(", txt)
    if not m:
        raise ValueError(path + ": not an ASCII list")
    body = txt[m.end(): txt.rindex(")")]
    return np.array(body.replace("(", " ").replace(")", " ").split(), float)


def compare(dir_a, dir_b, rtol):
    ok = True
    for blk in BLOCKS:
        errs, scale, seen = [], 0.0, False
        for part in ("diag", "upper", "lower"):
            pa, pb = os.path.join(dir_a, f"{blk}_{part}"), os.path.join(dir_b, f"{blk}_{part}")
            if not (os.path.exists(pa) and os.path.exists(pb)):
                continue
            a, b = read_list(pa), read_list(pb)
            if a.shape != b.shape:
                print(f"{blk:>10s} {part}: size {a.size} vs {b.size}  FAILED")
                ok = False
                continue
            seen = True
            if a.size:
                errs.append(np.abs(a - b).max())
                scale = max(scale, np.abs(a).max(), np.abs(b).max())
        if not seen:
            print(f"{blk:>10s}  not in both dumps")
            continue
        rel = (max(errs) if errs else 0.0) / (scale or 1.0)
        good = rel <= rtol
        ok &= good
        print(f"{blk:>10s}  max rel err {rel:.3e}  (bar {rtol:.1e})  {'ok' if good else 'FAILED'}")
    return ok


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dir_a")
    ap.add_argument("dir_b")
    ap.add_argument("--rtol", type=float, default=1e-12)
    a = ap.parse_args(argv)
    return 0 if compare(a.dir_a, a.dir_b, a.rtol) else 1


if __name__ == "__main__":
    sys.exit(main())
