"""Host-side parsing of the standalone driver dbnsB200 (no GPU needed: -parseOnly stops before the device is touched):
OpenFOAM dictionaries of the shipped VKI-LS89 tutorial with `nonuniform List<...>` entries spliced in — the internalField of a
written time directory and a total-pressure profile on the inlet."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "icsfoam_b200", "host", "dbnsB200")
MESHLIB = os.path.join(ROOT, "icsfoam_b200", "meshtools", "libicsmesh.so")
REF = "/root/reference/tutorials/VKI-LS89"


def run(case_dir, *flags):
    return subprocess.run([DRIVER, case_dir, *flags], env=dict(os.environ, ICSMESH_LIB=MESHLIB), capture_output=True, text=True, timeout=300)


def foam_list(values):
    values = np.asarray(values)
    if values.ndim == 1:
        body = "\n".join(repr(float(v)) for v in values)
        return f"nonuniform List<scalar> {len(values)}\n(\n{body}\n)"
    body = "\n".join("(" + " ".join(repr(float(x)) for x in v) + ")" for v in values)
    return f"nonuniform List<vector> {len(values)}\n(\n{body}\n)"


@pytest.mark.skipif(not (os.path.isdir(REF) and os.path.exists(DRIVER)), reason="reference tutorial or driver binary not present")
def test_driver_parses_nonuniform_fields(tmp_path):
    case = tmp_path / "vki"
    shutil.copytree(REF, case)
    r = run(str(case), "-parseOnly")
    assert r.returncode == 0 and "parse ok: p in [100000, 100000], 0 non-uniform patch entries" in r.stdout, r.stdout + r.stderr
    N = 28059
    pvals = 1e5 + np.arange(N) * 0.5
    ptxt = (case / "0" / "p").read_text()
    ptxt = ptxt.replace("internalField   uniform 1e5;", "internalField   " + foam_list(pvals) + ";")
    ptxt = re.sub(r"p0\s+uniform 160500;", "p0 " + foam_list(160500.0 + np.arange(105)) + ";", ptxt)
    (case / "0" / "p").write_text(ptxt)
    utxt = (case / "0" / "U").read_text().replace("internalField   uniform (100 0 0);", "internalField   " + foam_list(np.tile([100.0, 1.0, 0.0], (N, 1))) + ";")
    (case / "0" / "U").write_text(utxt)
    r = run(str(case), "-parseOnly")
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"parse ok: p in [100000, {1e5 + (N - 1) * 0.5:g}], 1 non-uniform patch entries" in r.stdout, r.stdout
    # a list of the wrong length is an error, not a silent truncation
    (case / "0" / "T").write_text((case / "0" / "T").read_text().replace("internalField   uniform 400;", "internalField   " + foam_list(np.full(7, 400.0)) + ";"))
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "nonuniform internalField of T" in r.stderr


@pytest.mark.skipif(not (os.path.isdir(REF) and os.path.exists(DRIVER)), reason="reference tutorial or driver binary not present")
def test_driver_refuses_an_implicit_tref_and_reads_the_smooth_solver(tmp_path):
    """hConst without `Tref` means Tstd in OpenFOAM v2112 (e = Cv T - Cp Tstd): refused, not silently read as 0; the solver
    dictionary of smoothSolverCoupled (smoother Jacobi; nSweeps) is accepted, an unknown smoother is a FatalError."""
    case = tmp_path / "vki"
    shutil.copytree(REF, case)
    thermo = case / "constant" / "thermophysicalProperties"
    txt = thermo.read_text()
    assert re.search(r"Tref\s+0;", txt)
    thermo.write_text(re.sub(r"Tref\s+0;", "", txt))
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "Tref 0" in r.stderr, r.stdout + r.stderr
    thermo.write_text(txt)
    sol = case / "system" / "fvSolution"
    stxt = sol.read_text()
    m = re.search(r"flowSolver\s*\{", stxt)
    depth, i = 1, m.end()
    while depth:
        depth += (stxt[i] == "{") - (stxt[i] == "}")
        i += 1
    new = "flowSolver\n{\n    solver smoothSolverCoupled;\n    smoothSolverCoupled\n    {\n        smoother %s;\n        nSweeps 2;\n        maxIter 8;\n        tolerance 1e-12;\n        relTol 1e-3;\n    }\n}\n"
    sol.write_text(stxt[:m.start()] + new % "Jacobi" + stxt[i:])
    r = run(str(case), "-parseOnly")
    assert r.returncode == 0 and "parse ok" in r.stdout, r.stdout + r.stderr
    sol.write_text(stxt[:m.start()] + new % "GaussSeidel" + stxt[i:])
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "Unknown smoother type GaussSeidel" in r.stderr, r.stdout + r.stderr


@pytest.mark.skipif(not (os.path.isdir(REF) and os.path.exists(DRIVER)), reason="reference tutorial or driver binary not present")
def test_driver_refuses_what_it_does_not_implement(tmp_path):
    """Settings the driver does not implement are FatalErrors, not silent substitutions: an inner ddt scheme other than
    steadyState / Euler / backward, pseudoTime/resetPseudo, adjustTimeStep; a freestreamPressure patch takes its velocity from the
    freestreamValue of the U patch field on the same patch (macros resolved), and needs that patch field to be `freestream`."""
    case = tmp_path / "vki"
    shutil.copytree(REF, case)

    def edit(rel, old, new, count=1):
        path = case / rel
        txt = path.read_text()
        assert len(re.findall(old, txt)) >= 1, (rel, old)
        path.write_text(re.sub(old, new, txt, count=count))
        return txt

    keep = edit("system/fvSchemes", r"default dualTime rPseudoDeltaT steadyState;", "default dualTime rPseudoDeltaT CrankNicolson 0.9;")
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "inner ddt scheme" in r.stderr and "not supported" in r.stderr, r.stdout + r.stderr
    (case / "system" / "fvSchemes").write_text(keep)

    keep = edit("system/fvSolution", r"nPseudoCorr\s+1;", "nPseudoCorr  1;\n    resetPseudo true;")
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "resetPseudo" in r.stderr, r.stdout + r.stderr
    (case / "system" / "fvSolution").write_text(keep)

    keep = edit("system/controlDict", r"adjustTimeStep\s+no;", "adjustTimeStep  yes;")
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "adjustTimeStep" in r.stderr, r.stdout + r.stderr
    (case / "system" / "controlDict").write_text(keep)

    # freestreamPressure on the outlet while U stays pressureInletOutletVelocity: refused
    keep_p = edit("0/p", r"outlet\s*\{\s*type\s+fixedValue;\s*value\s+uniform 82000;", "outlet\n    {\n        type freestreamPressure;\n        freestreamValue uniform 82000;")
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "freestreamPressure needs a freestream U patch field" in r.stderr, r.stdout + r.stderr
    # ... with a freestream U patch whose freestreamValue is a macro of the file: accepted
    edit("0/U", r"outlet\s*\{\s*type\s+pressureInletOutletVelocity;\s*tangentialVelocity\s+uniform \(0 0 0\);", "outlet\n    {\n         type freestream;\n         freestreamValue $internalField;")
    r = run(str(case), "-parseOnly")
    assert r.returncode == 0 and "parse ok" in r.stdout, r.stdout + r.stderr
    # ... supersonic true: refused
    (case / "0" / "p").write_text((case / "0" / "p").read_text().replace("freestreamValue uniform 82000;", "freestreamValue uniform 82000;\n        supersonic true;"))
    r = run(str(case), "-parseOnly")
    assert r.returncode != 0 and "supersonic" in r.stderr, r.stdout + r.stderr
    (case / "0" / "p").write_text(keep_p)
