"""bench.py's output contract on the arm that runs without a GPU (--impl reference): one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-cells-per-dim", "12"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("Mcell-iterations/s") and d["unit"] == "Mcell-it/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_both_arms_share_one_config_object_and_the_cpu_arm_uses_every_core():
    """The driver compares the `config` of the GPU line with that of the reference line: both come from Workload.config; the
    sample the CPU arm really ran is stated in cpu_baseline.sample / run.sample_cells; partitions = all host cores."""
    import argparse
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.factor3(16) == (4, 2, 2) and bench.factor3(13) == (13, 1, 1) and bench.factor3(64) == (4, 4, 4) and bench.factor3(48) == (4, 4, 3)
    assert bench.host_threads() >= 1
    args = argparse.Namespace(workload="onera344", n=344, n_cpu=12, bump_nx=1280, bump_ny=1040, bump_nx_cpu=40, bump_ny_cpu=30)
    want = bench.Workload(args).config
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-cells-per-dim", "12"],
                       capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["config"] == want and "344^3" in d["config"]["workload"] and d["config"]["cells"] == 344 ** 3
    assert d["run"]["sample_cells"] == 12 ** 3 and "12^3" in d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["cores"] == bench.host_threads()
    bump = bench.Workload(argparse.Namespace(workload="bump4m", n=344, n_cpu=12, bump_nx=1280, bump_ny=1040, bump_nx_cpu=40, bump_ny_cpu=30))
    assert bump.config["cells"] == 3 * 1280 * 1040 and "Minmod" in bump.config["workload"]


def test_reference_arm_runs_the_harmonic_balance_workload():
    """--workload vki-hb (C5 ii): the CPU arm is the reference-structured HB oracle world on the shipped VKI-LS89 mesh."""
    import pytest
    from icsfoam_b200 import cases
    if cases.tutorial_dir("VKI-LS89") is None:
        pytest.skip("VKI-LS89 tutorial not found")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "vki-hb"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["config"]["cells"] == 3 * 28059 and "Harmonic Balance" in d["config"]["workload"] and d["value"] > 0
    assert d["run"]["sample_cells"] == 3 * 28059 and d["run"]["restarts_per_step"] >= 1
