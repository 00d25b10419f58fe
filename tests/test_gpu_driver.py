"""The standalone driver icsfoam_b200/host/dbnsB200 (dbnsFoam's loop over the C ABI, reading OpenFOAM case directories
verbatim) against the Python host path on the two tutorials whose meshes the reference ships: it must read the same
case out of the dictionaries (0/p,U,T incl. `$internalField` macros, fvSchemes, fvSolution, thermophysicalProperties,
polyMesh incl. the cyclic pair) that icsfoam_b200.cases builds by hand, i.e. print the same residual history.  GPU only."""
import os
import re
import subprocess

import numpy as np
import pytest

from icsfoam_b200 import cases

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "icsfoam_b200", "host", "dbnsB200")
MESHLIB = os.path.join(ROOT, "icsfoam_b200", "meshtools", "libicsmesh.so")


def run_driver(case_dir, steps):
    r = subprocess.run([DRIVER, case_dir, "-maxSteps", str(steps)], env=dict(os.environ, ICSMESH_LIB=MESHLIB), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.rstrip().endswith("End")
    init = [[float(x) for x in re.findall(r"[-+0-9.eE]+", ln.split("=", 1)[1])] for ln in r.stdout.splitlines() if ln.startswith("Initial residual")]
    its = [int(ln.split()[-1]) for ln in r.stdout.splitlines() if ln.startswith("No Iterations")]
    return np.array(init), its


def staged(name):
    return cases.tutorial_dir(name) or os.path.join(ROOT, "cases_local", name)


@pytest.mark.skipif(not os.path.isdir(staged("VKI-LS89") + "/system"), reason="VKI-LS89 tutorial not staged ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_driver_runs_vki_ls89_tutorial_directory(gpu_context):
    init, its = run_driver(staged("VKI-LS89"), 6)
    case = cases.vki_ls89(staged("VKI-LS89") + "/constant/polyMesh")
    g = case.apply(gpu_context())
    for k in range(6):
        r = g.iterate(case.controls)
        assert its[k] == r.n_iterations, k
        want = np.array(list(r.s_init) + list(r.v_init))
        assert np.abs(init[k][:4] - want[:4]).max() <= 1e-5 * np.abs(want[:4]).max(), (k, init[k], want)   # 6 printed digits


@pytest.mark.skipif(not os.path.isdir(staged("forwardStep") + "/system"), reason="forwardStep tutorial not staged ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_driver_runs_forward_step_tutorial_directory(gpu_context):
    init, its = run_driver(staged("forwardStep"), 1)     # one physical time step = up to nPseudoCorr pseudo iterations
    case = cases.forward_step(staged("forwardStep") + "/constant/polyMesh")
    g = case.apply(gpu_context())
    g.new_time_step()
    assert len(its) >= 3
    for k in range(min(len(its), 5)):
        r = g.iterate(case.controls)
        assert its[k] == r.n_iterations, k
        want = np.array(list(r.s_init) + list(r.v_init))
        assert np.abs(init[k][:4] - want[:4]).max() <= 1e-5 * np.abs(want[:4]).max(), (k, init[k], want)


@pytest.mark.skipif(not os.path.isdir(staged("VKI-LS89") + "/system"), reason="VKI-LS89 tutorial not staged ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_driver_writes_a_time_directory_that_it_can_read_back(gpu_context, tmp_path):
    """-writeFields: p, U, T of the last step as <case>/<time>/{p,U,T}; the written internalField lists equal the state of the
    Python host path after the same 3 iterations to the 17 printed digits, and the files parse as `nonuniform List<...>` fields
    when spliced into 0/ (restart)."""
    import shutil
    case_dir = str(tmp_path / "vki")
    shutil.copytree(staged("VKI-LS89"), case_dir)
    r = subprocess.run([DRIVER, case_dir, "-maxSteps", "3", "-writeFields"], env=dict(os.environ, ICSMESH_LIB=MESHLIB), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fields written to" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    case = cases.vki_ls89(staged("VKI-LS89") + "/constant/polyMesh")
    g = case.apply(gpu_context())
    for _ in range(3):
        g.iterate(case.controls)
    st = g.state_get()

    def internal(path, nc):
        txt = open(path).read()
        body = txt[txt.index("internalField"):txt.index("boundaryField")]
        nums = np.array([float(x) for x in re.findall(r"[-+]?\d[\d.]*(?:[eE][-+]?\d+)?", body.split("\n(\n", 1)[1])])
        return nums.reshape(-1, nc) if nc > 1 else nums

    assert np.array_equal(internal(os.path.join(case_dir, "3", "p"), 1), st["p"])
    assert np.array_equal(internal(os.path.join(case_dir, "3", "T"), 1), st["T"])
    assert np.array_equal(internal(os.path.join(case_dir, "3", "U"), 3), st["U"])
    # restart: splice the written internalField into 0/p and let the driver parse it
    ptxt = open(os.path.join(case_dir, "3", "p")).read()
    written = ptxt[ptxt.index("internalField"):ptxt.index("boundaryField")]
    p0 = open(os.path.join(case_dir, "0", "p")).read().replace("internalField   uniform 1e5;", written)
    open(os.path.join(case_dir, "0", "p"), "w").write(p0)
    r = subprocess.run([DRIVER, case_dir, "-parseOnly"], env=dict(os.environ, ICSMESH_LIB=MESHLIB), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "parse ok" in r.stdout, r.stdout + r.stderr
    lo, hi = [float(x) for x in re.search(r"p in \[([^,]+), ([^\]]+)\]", r.stdout).groups()]
    assert abs(lo - st["p"].min()) <= 1e-5 * st["p"].min() and abs(hi - st["p"].max()) <= 1e-5 * st["p"].max()
