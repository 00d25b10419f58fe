"""Harmonic Balance on the CPU: the host-side HB operators (icsfoam_b200/hb.py) against their known
answers, the replicated-mesh layout, and the oracle's global (2 nO, nO) system against a dense numpy
assembly of the same blocks."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases, hb
from oracle.pyoracle import HB, Oracle


def test_omega_list_order():
    ol = hb.omega_list([3.0, -5.0], [2, 1])
    assert np.array_equal(ol, [0.0, 3.0, 6.0, 5.0, -5.0, -6.0, -3.0])
    with pytest.raises(ValueError):
        hb.omega_list([1.0], [1, 2])


def test_D_one_harmonic_three_instants_is_skew_circulant():
    om = 2 * np.pi * 40.0
    snaps, (D,) = hb.set_instants([hb.omega_list([om], [1])], 3, selected_period=2 * np.pi / om)
    assert np.allclose(snaps, np.arange(3) * (2 * np.pi / om) / 3)
    k = om / np.sqrt(3.0)
    assert np.allclose(D, [[0, k, -k], [-k, 0, k], [k, -k, 0]], atol=1e-12 * om)
    # exact spectral derivative of the resolved harmonic, zero on constants
    assert np.allclose(D @ np.sin(om * snaps), om * np.cos(om * snaps), atol=1e-11 * om)
    assert np.allclose(D @ np.cos(om * snaps), -om * np.sin(om * snaps), atol=1e-11 * om)
    assert np.allclose(D @ np.ones(3), 0, atol=1e-12 * om)


def test_D_two_harmonics_five_instants():
    om = 7.5
    snaps, (D,) = hb.set_instants([hb.omega_list([om], [2])], 5, selected_period=2 * np.pi / om)
    for k in (1, 2):
        assert np.allclose(D @ np.sin(k * om * snaps), k * om * np.cos(k * om * snaps), atol=1e-11 * om)
    assert np.allclose(D, -D.T, atol=1e-12 * om)


def test_condition_number_search_picks_uniform_period_for_one_frequency():
    om = 11.0
    ol = hb.omega_list([om], [1])
    snaps, _ = hb.set_instants([ol], 3)                      # no selectedPeriod: search over [T0, 5 T0]
    assert hb.condition_number(snaps, ol) < 1.0 + 1e-6       # uniform sampling of one period is perfectly conditioned
    assert hb.condition_number(np.array([0.0, 0.01, 0.02]) / om, ol) > 10.0
    with pytest.raises(ValueError):
        hb.set_instants([ol], 4)                             # instantsNumber must match unless oversampling
    s2, (D2,) = hb.set_instants([ol], 5, oversampling=True)
    assert D2.shape == (5, 5)


def test_two_frequencies_non_harmonic():
    ol = hb.omega_list([10.0, 13.0], [1, 1])
    snaps, (D,) = hb.set_instants([ol], 5)
    assert hb.condition_number(snaps, ol) < 4.0
    for w in (10.0, 13.0):
        assert np.allclose(D @ np.sin(w * snaps), w * np.cos(w * snaps), atol=1e-9 * w)


def test_replicated_mesh_layout():
    base = cases.periodic_box(4).mesh
    m = hb.replicate(base, 3)
    N, F, FT = base.n_cells, base.n_internal_faces, base.n_faces
    assert (m.n_cells, m.n_internal_faces, m.n_faces) == (3 * N, 3 * F, 3 * FT)
    assert np.all(m.owner[:3 * F] < m.neighbour)
    for K in range(3):
        assert np.array_equal(m.owner[K * F:(K + 1) * F], base.owner[:F] + K * N)
        assert np.array_equal(m.neighbour[K * F:(K + 1) * F], base.neighbour + K * N)
    for K in range(3):
        for i, p in enumerate(base.patches):
            q = m.patches[K * len(base.patches) + i]
            assert q["size"] == p["size"] and q["name"] == f"{p['name']}@{K}"
            f = np.arange(q["start"], q["start"] + q["size"])
            assert np.array_equal(m.owner[f], base.owner[p["start"]:p["start"] + p["size"]] + K * N)
            if p["kind"] == capi.CYCLIC:
                assert q["nbr_patch"] == p["nbr_patch"] + K * len(base.patches)
    # the oracle accepts it as an ordinary mesh
    Oracle().mesh_set(m)


def dense_global(H, case):
    """(5 nO N)^2 dense matrix of the global system from the instances' LDU blocks + V D[J][K] coupling.
    Row/column order: instance-major, per cell (rho, rhoUx, rhoUy, rhoUz, rhoE)."""
    nO, N = H.n, H.N
    mesh = case.base.mesh
    F = mesh.n_internal_faces
    own, nei = mesh.owner[:F], mesh.neighbour
    A = np.zeros((nO * N * 5, nO * N * 5))
    # block id -> (row comps, col comps) in the 5-block
    comp = {0: ([0], [0]), 1: ([0], [4]), 2: ([4], [0]), 3: ([4], [4]), 4: ([0], [1, 2, 3]), 5: ([4], [1, 2, 3]),
            6: ([1, 2, 3], [0]), 7: ([1, 2, 3], [4]), 8: ([1, 2, 3], [1, 2, 3])}
    for K, o in enumerate(H.inst):
        base = K * N * 5
        for b in range(9):
            d, u, l = o.matrix_get_ldu(b)
            rows, cols = comp[b]
            nr, ncol = len(rows), len(cols)
            d, u, l = d.reshape(N, nr, ncol), u.reshape(F, nr, ncol), l.reshape(F, nr, ncol)
            for i, r in enumerate(rows):
                for j, c in enumerate(cols):
                    A[base + np.arange(N) * 5 + r, base + np.arange(N) * 5 + c] += d[:, i, j]
                    A[base + nei * 5 + r, base + own * 5 + c] += l[:, i, j]
                    A[base + own * 5 + r, base + nei * 5 + c] += u[:, i, j]
    zone = case.zone_of_cell
    for J in range(nO):
        for K in range(nO):
            if J == K:
                continue
            for cell in range(N):
                if zone is not None and zone[cell] < 0:
                    continue
                v = mesh.V[cell] * case.D[0][J, K]
                for r in range(5):
                    A[(J * N + cell) * 5 + r, (K * N + cell) * 5 + r] += v
    return A


def pack(a, b, c):
    return np.concatenate([a[:, None], b, c[:, None]], axis=1).reshape(-1)


@pytest.mark.parametrize("zoned", [False, True])
def test_oracle_global_system_against_dense(zoned):
    # no cyclic pair here: the dense assembly above has no interface coefficients
    case = cases.hb_box(4, 3, zoned=zoned, cyclic=False)
    H = HB(case)
    H.assemble()
    A = dense_global(H, case)
    rng = np.random.default_rng(3)
    NT = case.mesh.n_cells
    x = (rng.standard_normal(NT), rng.standard_normal((NT, 3)), rng.standard_normal(NT))
    y = pack(*H.matrix_mul(*x))
    assert not any(p["kind"] == capi.CYCLIC for p in case.base.mesh.patches)
    assert np.allclose(y, A @ pack(*x), rtol=1e-12, atol=1e-12 * np.abs(y).max())
    # preconditioned GMRES reaches the dense solution of the same system
    ctl = capi.solver_controls("LUSGS", n_directions=10, max_iter=200, tolerance=1e-13, rel_tol=1e-11)
    (dr, dru, dre), res = H.solve_delta(ctl)
    b = pack(*H.residual())
    xs = np.linalg.solve(A, b)
    got = pack(dr, dru, dre)
    assert np.abs(got - xs).max() <= 1e-7 * np.abs(xs).max()
    ctl = capi.solver_controls("Jacobi", n_directions=10, max_iter=300, tolerance=1e-13, rel_tol=1e-11)
    (dr, dru, dre), res = H.solve_delta(ctl)
    assert np.abs(pack(dr, dru, dre) - xs).max() <= 1e-7 * np.abs(xs).max()
    # the HB source: S_J = -V sum_K D[J][K] W_K on the zone cells
    st = H.state_get()
    N = H.N
    for name, key in (("rho", 0), ("rhoE", 2)):
        W = st[name].reshape(3, N)
        S = -(case.D[0] @ W) * case.base.mesh.V[None, :]
        if case.zone_of_cell is not None:
            S[:, case.zone_of_cell < 0] = 0
        assert np.allclose(H.sources()[key].reshape(3, N), S, rtol=1e-12, atol=1e-12 * np.abs(S).max())


def test_oracle_cylindrical_source_equals_cartesian_for_axis_aligned_flow():
    """With every instance mesh identical the cylindrical decomposition is an orthonormal change of basis at a fixed
    point, so the cylindrical momentum source must equal the Cartesian one up to rounding (HBZone.C:521-651)."""
    a = HB(cases.hb_box(4, 3, cyl=False))
    b = HB(cases.hb_box(4, 3, cyl=True))
    sa, sb = a.sources()[1], b.sources()[1]
    assert np.allclose(sa, sb, rtol=1e-11, atol=1e-11 * np.abs(sa).max())
    assert not np.array_equal(sa, sb)


def test_oracle_hb_iteration_reduces_to_steady_for_identical_instances():
    """Identical instances: D annihilates constants in time, so every instance must follow the single-instance steady
    iteration (the HB source only adds rounding noise)."""
    case = cases.hb_box(4, 3)
    case.instances = [case.instances[0]] * 3
    c0 = case.instances[0]
    c0.schemes.pseudo_co_num_min = c0.schemes.pseudo_co_num   # keep Co fixed through the reference's SER quirk
    c0.schemes.pseudo_co_num_max = c0.schemes.pseudo_co_num
    H = HB(case)
    single = c0.apply(Oracle())
    for _ in range(3):
        H.iterate(case.controls)
        single.iterate(case.controls)
    ref = single.state_get()
    for K in range(3):
        for k in ("rho", "rhoU", "rhoE"):
            got = H.inst[K].state_get()[k]
            assert np.abs(got - ref[k]).max() <= 1e-9 * np.abs(ref[k]).max(), (K, k)


def test_oracle_hb_converges_and_responds_to_unsteady_inlet():
    case = cases.hb_box(5, 3, co=20.0)
    H = HB(case)
    first = None
    for it in range(40):
        res = H.iterate(case.controls)
        first = first if first is not None else res["s_init"].copy()
    assert np.all(res["s_init"] < 0.2 * first)
    st = [o.state_get()["p"] for o in H.inst]
    # the instances differ (the inlet state oscillates) ...
    assert np.abs(st[0] - st[1]).max() > 1e-4 * np.abs(st[0]).max()


def test_cpp_host_mirror_HBZone_matches_python(tmp_path):
    """The C++ host mirror (icsfoam_b200/host/icsfoamB200.H: HBZone / HBZoneList) computes the same snapshots and D."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "hbz.cpp"
    src.write_text(r'''
#include <cstdio>
#include "icsfoamB200.H"
using namespace icsfoamB200;
int main()
{
    HBZoneList l;
    l.zones.resize(2);
    l.zones[0].setOmegaList({10.0, 13.0}, {1, 1});
    l.zones[1].setOmegaList({10.0}, {2});
    l.setInstants(5);
    for (double t : l.selectedSnapshots) std::printf("%.17g ", t);
    std::printf("\n");
    for (auto& z : l.zones) { for (double d : z.D) std::printf("%.17g ", d); std::printf("\n"); }
    HBZoneList u;
    u.zones.resize(1);
    u.zones[0].setOmegaList({251.0}, {1});
    u.setInstants(3, false, 2 * M_PI / 251.0);
    for (double d : u.zones[0].D) std::printf("%.17g ", d);
    std::printf("\n");
    return 0;
}
''')
    exe = tmp_path / "hbz"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), "-I", os.path.join(root, "icsfoam_b200", "host"),
                    str(src), "-o", str(exe), "-Wl,--unresolved-symbols=ignore-all"], check=True)  # no icsb200_* call is made
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.strip().split("\n")
    rows = [np.array([float(x) for x in line.split()]) for line in out]
    ols = [hb.omega_list([10.0, 13.0], [1, 1]), hb.omega_list([10.0], [2])]
    snaps, Ds = hb.set_instants(ols, 5)
    assert np.allclose(rows[0], snaps, rtol=1e-12)
    assert np.allclose(rows[1].reshape(5, 5), Ds[0], rtol=1e-9, atol=1e-9)
    assert np.allclose(rows[2].reshape(5, 5), Ds[1], rtol=1e-9, atol=1e-9)
    _, (D3,) = hb.set_instants([hb.omega_list([251.0], [1])], 3, selected_period=2 * np.pi / 251.0)
    assert np.allclose(rows[3].reshape(3, 3), D3, rtol=1e-10, atol=1e-10 * 251)


def test_hb_oracle_reproduces_golden_fixture():
    """Drift pin of the reference-structured HB oracle (tests/golden/make_golden_hb.py)."""
    import os
    from tests.common import GOLDEN
    from tests.golden import make_golden_hb as mg
    case = mg.make_case()
    got = mg.run(HB(case), case, case.mesh.n_cells)
    gold = np.load(os.path.join(GOLDEN, "hb_box_roe_3instants.npz"))
    for k in gold.files:
        assert np.array_equal(got[k], gold[k]), k


def test_partitioned_hb_oracle_matches_single_domain():
    """Multi-rank Harmonic Balance: every rank holds all instances of its partition (HB instants are not sharded).
    With block-Jacobi preconditioning the P-rank HB world reproduces the single-domain HB run."""
    from icsfoam_b200 import capi
    from oracle.pyoracle import HBWorld
    case = cases.hb_box(5, 3, flux="ROE", cyclic=False, seed=3)
    ctl = capi.solver_controls("Jacobi", n_directions=5, max_iter=30, tolerance=1e-14, rel_tol=1e-9)
    H = HB(case)
    parts = case.partition(2, "x")
    W = HBWorld(parts)
    for it in range(3):
        r1, r2 = H.iterate(ctl), W.iterate(ctl)
        assert r1["n_iterations"] == r2[0]["n_iterations"] == r2[1]["n_iterations"]
        assert np.abs(r1["s_init"] - r2[0]["s_init"]).max() <= 1e-10
    s1 = H.state_get()
    N = case.base.mesh.n_cells
    for r, hc in enumerate(parts):
        s2 = W.ranks[r].state_get()
        idx = np.concatenate([K * N + hc.base.mesh.cell_global for K in range(3)])
        for k in ("rho", "rhoU", "rhoE"):
            assert np.abs(s2[k] - s1[k][idx]).max() <= 1e-9 * np.abs(s1[k]).max(), (r, k)


def test_phase_lag_cyclic_travelling_wave():
    """phaseLagCyclic (phaseLagCyclicFvPatchField.C:160-398): D_pl = Re(EInv M(IBPA) E) shifts a single-harmonic time signal by the
    inter-blade phase angle, so with p_K(cell) = p0 + a cos(omega t_K + phi_cell) in every time instance the value a face sees
    across the pair is p0 + a cos(omega t_K + phi_neighbour +- IBPA) — owner side +, neighbour side - — to rounding; IBPA = 0
    gives back the plain cyclic pair; the run stays finite."""
    om = 2 * np.pi * 40.0
    omegas = hb.omega_list([om], [1])
    assert np.allclose(hb.phase_lag_operator(np.arange(3) / 3 * (2 * np.pi / om), omegas, 0.0), np.eye(3), atol=1e-14)
    case = cases.hb_box(5, 3, omega=om, flux="ROE", seed=3)
    ibpa = 0.7
    N = case.base.mesh.n_cells
    phi = np.random.default_rng(0).random(N) * 2 * np.pi
    for c, t in zip(case.instances, case.snapshots):
        c.p = 1e5 + 3e3 * np.cos(om * t + phi)
        c.U = np.stack([50 + 10 * np.cos(om * t + phi + 0.3), 5 * np.cos(om * t + phi), 0 * phi], 1)
    case.p = np.concatenate([c.p for c in case.instances])
    case.U = np.concatenate([c.U for c in case.instances])
    case.with_phase_lag("xmin", "xmax", ibpa, omegas)
    H = HB(case)
    m = case.base.mesh
    F = m.n_internal_faces
    fa, fb = m.patch_faces("xmin"), m.patch_faces("xmax")
    for K, t in enumerate(case.snapshots):
        b = H.inst[K].boundary_get()
        assert np.abs(b["p"][fa - F] - (1e5 + 3e3 * np.cos(om * t + phi[m.owner[fb]] + ibpa))).max() <= 1e-12 * 1e5
        assert np.abs(b["p"][fb - F] - (1e5 + 3e3 * np.cos(om * t + phi[m.owner[fa]] - ibpa))).max() <= 1e-12 * 1e5
        assert np.abs(b["U"][fa - F, 0] - (50 + 10 * np.cos(om * t + phi[m.owner[fb]] + 0.3 + ibpa))).max() <= 1e-12 * 60
        assert np.array_equal(b["T"][fa - F], H.inst[K].state_get()["T"][m.owner[fb]])     # T is not a phase-lagged field
    r = H.iterate(case.controls, 3)
    assert np.isfinite(H.state_get()["rho"]).all() and r["s_init"].max() < 1.0


def test_cylindrical_momentum_source_second_reading():
    """HBZone::addSource for vectors with cylCoords (HBZone.C:521-651) read a second time in numpy: every instance's momentum is
    decomposed along (rHat, axisHat x rHat, axisHat) at ITS cell centre, mixed by D, and the sum is turned back into Cartesian
    components at the centre of the receiving instance."""
    case = cases.hb_box(4, 3, cyl=True)
    H = HB(case)
    got = H.sources()[1].reshape(3, H.N, 3)
    st = H.state_get()
    W = st["rhoU"].reshape(3, H.N, 3)
    V = case.base.mesh.V
    D = case.D[0]
    axis = np.array(case.rotation_axis[:3], float)
    a = axis / np.linalg.norm(axis)
    centre = np.array(case.rotation_centre[:3], float)
    C = case.base.mesh.C                                     # the instance meshes are copies: same centres
    r = C - centre
    r = r - (r @ a)[:, None] * a
    rhat = r / np.linalg.norm(r, axis=1)[:, None]
    that = np.cross(a, rhat)
    want = np.zeros_like(got)
    for J in range(3):
        cyl = np.zeros((H.N, 3))
        for K in range(3):
            ucyl = np.column_stack([(W[K] * rhat).sum(1), (W[K] * that).sum(1), W[K] @ a])
            cyl += (V * D[J, K])[:, None] * ucyl
        want[J] = -(cyl[:, :1] * rhat + cyl[:, 1:2] * that + cyl[:, 2:3] * a)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
