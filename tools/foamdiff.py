#!/usr/bin/env python
"""Compare OpenFOAM ASCII field files (vol*/surface* Scalar / Vector / Tensor fields) entry by entry.

Closes the "parity unpinned" gap on a machine that has OpenFOAM v2112 + ICSFoam: run the
reference there, copy its time directory next to the one `dbnsB200 -writeFlux` wrote, and

    python tools/foamdiff.py <referenceCase>/<time> <b200Case>/<time> [--rtol 1e-8] [--fields p U T rho phi]

prints, per field, the largest error relative to the field's largest magnitude over the internal field and over every
boundary patch both files carry a `value` for, and exits non-zero when a bar is exceeded (1e-12 for the face fluxes of the
first iteration, 1e-8 for converged fields — the north-star tolerances).  Recipe: tools/openfoam_golden/README.md.
"""
import argparse
import os
import re
import sys

import numpy as np

_NCOMP = {"scalar": 1, "vector": 3, "symmTensor": 6, "tensor": 9}


def _strip(txt):
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return re.sub(r"//[^\n]*", "", txt)


def _entry(txt, pos):
    """value of a field entry starting at `pos` (just after the keyword): (array or float-array of one row, is_uniform, end)"""
    m = re.compile(r"\s*uniform\s+").match(txt, pos)
    if m:
        end = txt.index(";", m.end())
        v = np.array(txt[m.end():end].replace("(", " ").replace(")", " ").split(), float)
        return v, True, end + 1
    m = re.compile(r"\s*nonuniform\s+List<(\w+)>\s*").match(txt, pos)
    if not m:
        raise ValueError("unsupported entry: " + txt[pos:pos + 60].strip())
    nc = _NCOMP[m.group(1)]
    m2 = re.compile(r"(\d+)\s*\(").match(txt, m.end())
    if not m2:                                   # "nonuniform List<scalar> 0()" or "0;"
        end = txt.index(";", m.end())
        return np.zeros((0, nc)), False, end + 1
    n = int(m2.group(1))
    # the list body ends at the parenthesis that closes the outer list
    depth, i = 1, m2.end()
    start = i
    if nc == 1:
        i = txt.index(")", start)
    else:
        while depth:
            c = txt[i]
            depth += (c == "(") - (c == ")")
            i += 1
        i -= 1
    v = np.array(txt[start:i].replace("(", " ").replace(")", " ").split(), float).reshape(n, nc)
    end = txt.index(";", i)
    return v, False, end + 1


def read_field(path):
    """{'internal': (array, uniform), 'patches': {name: (array, uniform)}} — patches without a `value` entry are left out."""
    txt = _strip(open(path).read())
    h = txt.find("FoamFile")
    if h >= 0:
        txt = txt[txt.index("}", h) + 1:]
    m = re.search(r"\binternalField\b", txt)
    if not m:
        raise ValueError(path + ": no internalField")
    internal, uni, end = _entry(txt, m.end())
    out = {"internal": (internal, uni), "patches": {}}
    b = re.search(r"\bboundaryField\s*\{", txt[end:])
    if not b:
        return out
    i = end + b.end()
    pat = re.compile(r"\s*([\w\.\-\"\*\(\)\|]+)\s*\{")
    while True:
        m = pat.match(txt, i)
        if not m:
            break
        name, j, depth = m.group(1), m.end(), 1
        k = j
        while depth:
            depth += (txt[k] == "{") - (txt[k] == "}")
            k += 1
        body = txt[j:k - 1]
        v = re.search(r"(?<![\w])value\s", body)
        if v:
            try:
                arr, u, _ = _entry(body, v.end() - 1)
                out["patches"][name] = (arr, u)
            except ValueError:
                pass
        i = k
    return out


def _cmp(a, b):
    (va, ua), (vb, ub) = a, b
    va, vb = np.atleast_2d(va) if not ua else va.reshape(1, -1), np.atleast_2d(vb) if not ub else vb.reshape(1, -1)
    if va.shape[0] != vb.shape[0]:
        if va.shape[0] == 1:
            va = np.repeat(va, vb.shape[0], 0)
        elif vb.shape[0] == 1:
            vb = np.repeat(vb, va.shape[0], 0)
        else:
            return None
    if va.size == 0:
        return 0.0, 0.0
    return float(np.abs(va - vb).max()), float(max(np.abs(va).max(), np.abs(vb).max()))


def diff_fields(path_a, path_b):
    """(max relative error, {part: (abs err, scale)}); relative to the largest magnitude of the whole field."""
    fa, fb = read_field(path_a), read_field(path_b)
    parts = {"internalField": _cmp(fa["internal"], fb["internal"])}
    for name in fa["patches"]:
        if name in fb["patches"]:
            parts[name] = _cmp(fa["patches"][name], fb["patches"][name])
    if any(v is None for v in parts.values()):
        bad = [k for k, v in parts.items() if v is None]
        raise ValueError(f"size mismatch in {bad} between {path_a} and {path_b}")
    scale = max(s for _, s in parts.values()) or 1.0
    return max(e for e, _ in parts.values()) / scale, parts


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("dir_a")
    ap.add_argument("dir_b")
    ap.add_argument("--fields", nargs="*", default=None)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--flux-rtol", type=float, default=None, help="bar for phi / phiUp / phiEp (default: --rtol)")
    a = ap.parse_args(argv)
    fields = a.fields or sorted(f for f in os.listdir(a.dir_a) if os.path.isfile(os.path.join(a.dir_a, f)) and os.path.isfile(os.path.join(a.dir_b, f)))
    worst_ok = True
    for f in fields:
        try:
            rel, parts = diff_fields(os.path.join(a.dir_a, f), os.path.join(a.dir_b, f))
        except (ValueError, KeyError) as e:
            print(f"{f:>10s}  skipped: {e}")
            continue
        bar = a.flux_rtol if (a.flux_rtol is not None and f in ("phi", "phiUp", "phiEp")) else a.rtol
        ok = rel <= bar
        worst_ok &= ok
        where = max(parts, key=lambda k: parts[k][0])
        print(f"{f:>10s}  max rel err {rel:.3e}  (bar {bar:.1e}, worst in {where}, {len(parts) - 1} patches)  {'ok' if ok else 'FAILED'}")
    return 0 if worst_ok else 1


if __name__ == "__main__":
    sys.exit(main())
