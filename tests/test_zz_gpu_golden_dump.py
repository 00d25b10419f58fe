"""`dbnsB200 -writeFlux`: the B200 side of the golden-data recipe (tools/openfoam_golden/README.md).  The time directory it
writes (p, U, T, rho, phi, phiUp, phiEp + eqSystem/) must hold what the Python host path computes for the same iterations, and
tools/foamdiff.py / compare_matrix.py must accept it against itself.  Runs last on purpose (new in this round's final session;
no GPU minutes were left to run it — DESIGN.md §6)."""
import importlib.util
import os
import shutil
import subprocess

import numpy as np
import pytest

from icsfoam_b200 import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "icsfoam_b200", "host", "dbnsB200")
MESHLIB = os.path.join(ROOT, "icsfoam_b200", "meshtools", "libicsmesh.so")
STAGED = cases.tutorial_dir("VKI-LS89") or os.path.join(ROOT, "cases_local", "VKI-LS89")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.skipif(not os.path.isdir(STAGED + "/system"), reason="VKI-LS89 tutorial not staged ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_write_flux_directory_matches_the_host_path(gpu_context, tmp_path):
    fd = _load(os.path.join(ROOT, "tools", "foamdiff.py"), "foamdiff")
    cm = _load(os.path.join(ROOT, "tools", "openfoam_golden", "compare_matrix.py"), "compare_matrix")
    case_dir = str(tmp_path / "vki")
    shutil.copytree(STAGED, case_dir)
    r = subprocess.run([DRIVER, case_dir, "-maxSteps", "2", "-writeFlux"], env=dict(os.environ, ICSMESH_LIB=MESHLIB), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fields written to" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    tdir = r.stdout.split("fields written to", 1)[1].split()[0]
    case = cases.vki_ls89(STAGED + "/constant/polyMesh")     # steady: one outer iteration per "time step" (same set-up as test_gpu_driver)
    g = case.apply(gpu_context())
    # the tutorial sets nPseudoCorr 1, for which pseudotimeControl::loop() prints no "pseudoTime: iteration" line (pseudotimeControl.C:229)
    n_pseudo = sum(1 for ln in r.stdout.splitlines() if ln.startswith("GMRES : Solving for"))
    assert n_pseudo == 2
    for it in range(n_pseudo):
        phi, phiUp, phiEp = g.calc_flux()            # as the driver does before every iteration; the last one is what dbnsFoam's phi holds at write time
        g.iterate(case.controls)
    st = g.state_get()
    F = case.mesh.n_internal_faces
    f = fd.read_field(os.path.join(tdir, "phi"))
    assert np.array_equal(f["internal"][0][:, 0], phi[:F])
    f = fd.read_field(os.path.join(tdir, "phiUp"))
    assert np.array_equal(f["internal"][0], phiUp[:F])
    f = fd.read_field(os.path.join(tdir, "phiEp"))
    assert np.array_equal(f["internal"][0][:, 0], phiEp[:F])
    assert np.array_equal(fd.read_field(os.path.join(tdir, "rho"))["internal"][0][:, 0], st["rho"])
    d, u, l = g.matrix_get_ldu(8)
    assert np.array_equal(cm.read_list(os.path.join(tdir, "eqSystem", "dVByV_0_0_diag")), d.reshape(-1))
    assert np.array_equal(cm.read_list(os.path.join(tdir, "eqSystem", "dVByV_0_0_upper")), u.reshape(-1))
    assert fd.main([tdir, tdir, "--rtol", "0"]) == 0
    assert cm.main([os.path.join(tdir, "eqSystem"), os.path.join(tdir, "eqSystem"), "--rtol", "0"]) == 0
