"""The planning tools under tools/ stay runnable: the LU-SGS cost model and the clustered-schedule prototype (CPU only)."""
import importlib.util
import os

import numpy as np

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


def _foam_field(path, cls, internal, patches):
    def lst(v):
        v = np.atleast_1d(v)
        if v.ndim == 1:
            return f"nonuniform List<scalar> {len(v)}\n(\n" + "\n".join(repr(float(x)) for x in v) + "\n)"
        return f"nonuniform List<vector> {len(v)}\n(\n" + "\n".join("(" + " ".join(repr(float(x)) for x in r) + ")" for r in v) + "\n)"
    with open(path, "w") as f:
        f.write("/*--- header ---*/\nFoamFile\n{\n    version 2.0;\n    format ascii;\n    class %s;\n    object x;\n}\n// comment\ndimensions [0 0 0 0 0 0 0];\n\n" % cls)
        f.write("internalField " + (internal if isinstance(internal, str) else lst(internal)) + ";\n\nboundaryField\n{\n")
        for name, body in patches.items():
            f.write(f"    {name}\n    {{\n")
            for k, v in body.items():
                f.write(f"        {k} " + (v if isinstance(v, str) else lst(v)) + ";\n")
            f.write("    }\n")
        f.write("}\n")


def test_foamdiff_reads_and_compares_openfoam_fields(tmp_path):
    """tools/foamdiff.py — the comparer of the golden-data recipe (tools/openfoam_golden/README.md)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("foamdiff", os.path.join(ROOT, "tools", "foamdiff.py"))
    fd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fd)
    rng = np.random.default_rng(0)
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    p = 1e5 * (1 + rng.random(50))
    U = rng.standard_normal((50, 3)) * 100
    for d, eps in ((a, 0.0), (b, 1e-10)):
        _foam_field(d / "p", "volScalarField", p * (1 + eps), {"inlet": {"type": "totalPressure", "p0": "uniform 101300", "gamma": "1.4", "value": p[:4] * (1 + eps)},
                                                                "walls": {"type": "zeroGradient"}, "frontAndBack": {"type": "empty"}})
        _foam_field(d / "U", "volVectorField", U, {"inlet": {"type": "fixedValue", "value": "uniform (1 2 3)" if eps == 0.0 else U[:4] * 0 + [1, 2, 3]},
                                                    "outlet": {"type": "inletOutlet", "inletValue": "uniform (0 0 0)", "value": U[:3]}})
        _foam_field(d / "T", "volScalarField", "uniform 300", {"inlet": {"type": "fixedValue", "value": "uniform 300"}})
    f = fd.read_field(str(a / "p"))
    assert np.array_equal(f["internal"][0][:, 0], p) and set(f["patches"]) == {"inlet"}
    fu = fd.read_field(str(a / "U"))
    assert np.array_equal(fu["internal"][0], U) and fu["patches"]["inlet"][1] and np.array_equal(fu["patches"]["outlet"][0], U[:3])
    rel, parts = fd.diff_fields(str(a / "p"), str(b / "p"))
    assert 0.4e-10 < rel < 1.1e-10 and set(parts) == {"internalField", "inlet"}
    assert fd.diff_fields(str(a / "U"), str(b / "U"))[0] == 0.0          # uniform value against the expanded list
    assert fd.main([str(a), str(b), "--rtol", "1e-8"]) == 0
    assert fd.main([str(a), str(b), "--rtol", "1e-12"]) == 1
    ref0 = "/root/reference/tutorials/VKI-LS89/0"
    if os.path.isdir(ref0):       # the reference's own field files parse (self-comparison)
        assert fd.main([ref0, ref0]) == 0


def test_compare_matrix_reads_openfoam_lists(tmp_path):
    """tools/openfoam_golden/compare_matrix.py: LDU dumps of the coupledMatrix sub-blocks (scalar / vector / tensor lists,
    uniform `N{v}` lists) against each other."""
    spec = importlib.util.spec_from_file_location("compare_matrix", os.path.join(ROOT, "tools", "openfoam_golden", "compare_matrix.py"))
    cm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cm)
    rng = np.random.default_rng(1)
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    for blk, nc in zip(cm.BLOCKS, (1, 1, 1, 1, 3, 3, 3, 3, 9)):
        for part, n in (("diag", 7), ("upper", 12), ("lower", 12)):
            v = rng.standard_normal((n, nc))
            for d, eps in ((a, 0.0), (b, 1e-13)):
                with open(d / f"{blk}_{part}", "w") as f:
                    rows = [repr(float(r[0])) if nc == 1 else "(" + " ".join(repr(float(x)) for x in r) + ")" for r in v * (1 + eps)]
                    f.write(f"\n{n}\n(\n" + "\n".join(rows) + "\n)\n\n")
    assert cm.read_list(str(a / "dVByV_0_0_diag")).shape == (63,)
    (a / "u").write_text("5{(1 2 3)}")
    assert np.array_equal(cm.read_list(str(a / "u")), np.tile([1.0, 2.0, 3.0], 5))
    assert cm.main([str(a), str(b)]) == 0
    assert cm.main([str(a), str(b), "--rtol", "1e-15"]) == 1


def test_general_cluster_schedule_on_ldu_addressing():
    """The clustered LU-SGS schedule built from owner / neighbour alone (structured and randomly renumbered meshes): intra-slab
    dependencies strictly descend in level, cross-slab ones are `lag` steps old, and the replay equals the sequential sweep."""
    from icsfoam_b200 import cases
    m = _load("lusgs_cluster_schedule")
    box = cases.onera_box(12).mesh
    F = box.n_internal_faces
    s = m.general_schedule(box.n_cells, box.owner[:F], box.neighbour, n_slabs=4, n_clusters=4, lag=2)
    r = m.check_general(box.n_cells, box.owner[:F], box.neighbour, s)
    assert r["global_levels"] == 3 * 12 - 2
    # 4 slabs of thickness 3: (12 + 12 + 3 - 2) levels per slab, each slab starts lag + 1 steps behind its predecessor
    assert list(s["depth"]) == [25] * 4 and list(s["off"]) == [0, 4, 8, 12] and r["chain"] == 25 + 12
    scr = cases.scrambled_box(6).mesh
    F = scr.n_internal_faces
    s = m.general_schedule(scr.n_cells, scr.owner[:F], scr.neighbour, n_slabs=6, n_clusters=3, lag=3)
    r = m.check_general(scr.n_cells, scr.owner[:F], scr.neighbour, s)
    assert r["chain"] >= r["global_levels"] // 2
