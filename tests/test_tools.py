"""The planning tools under tools/ stay runnable: the LU-SGS cost model and the clustered-schedule prototype (CPU only)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_clustered_lusgs_schedule_keeps_the_sequential_result():
    m = _load("lusgs_cluster_schedule")
    s, chain = m.check_and_replay(12, 4)        # asserts P1, P2 and bit-identity with the sequential sweep
    assert s == 3 and chain == 12 + 12 + 3 - 2 + 3 * 3


def test_lusgs_cost_model_reproduces_the_measured_sweeps():
    m = _load("lusgs_model")
    for n, meas in m.MEASURED_MS.items():
        assert abs(2e3 * m.sweep_today(n) - meas) < 0.2 * meas          # hop model within 20 % of the B200 measurements
        assert 2e3 * m.sweep_clustered(n)[0] <= 2e3 * m.sweep_today(n)
