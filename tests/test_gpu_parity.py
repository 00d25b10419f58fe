"""Parity of the CUDA path (through the C-ABI) with the CPU oracle and the golden fixtures.  GPU only (-m gpu).

Bars (BASELINE.json north_star): per-face fluxes and assembled Jacobian blocks 1e-12 relative (in practice bit-equal:
the kernels keep the reference's operand and accumulation order and are built with -fmad=false); GMRES increments and
fields after several outer iterations 1e-8 relative (reduction order differs from the CPU's sequential sums)."""
import os

import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from icsfoam_b200 import meshtools as mt
from oracle.pyoracle import Oracle
from tests.common import EXACT_KEYS, GOLDEN, SOLVE_KEYS, STATE_KEYS, TOL_EXACT, TOL_SOLVE, TOL_STATE, rel_err, run_sequence
from tests.golden.make_golden import CASES as GOLDEN_CASES

pytestmark = pytest.mark.gpu


def compare(out_gpu, out_ref):
    for k in EXACT_KEYS:
        assert rel_err(out_gpu[k], out_ref[k]) <= TOL_EXACT, k
    for k in SOLVE_KEYS:
        assert rel_err(out_gpu[k], out_ref[k]) <= TOL_SOLVE, k
    for k in STATE_KEYS:
        assert rel_err(out_gpu[k], out_ref[k]) <= TOL_STATE, k
    assert out_gpu["restarts"][0] == out_ref["restarts"][0]
    assert np.array_equal(out_gpu["history"][:, -1], out_ref["history"][:, -1])          # restarts per outer iteration
    assert rel_err(out_gpu["history"][:, :5], out_ref["history"][:, :5]) <= 1e-8           # residual history


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_against_golden_fixtures(name, gpu_context):
    case = GOLDEN_CASES[name]()
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = case.apply(gpu_context())
    out = run_sequence(g, case)
    compare(out, gold)
    assert g.launch_count() > 0


def FULLV(case):
    case.schemes.viscous_full_jacobian = 1
    return case


def TURB(x):
    """smooth positive muEff / alphaEff fields standing in for mu + mut, alpha + alphat of a turbulence model"""
    mu = 0.05 * (1.5 + np.sin(3.0 * x[:, 0]) * np.cos(2.0 * x[:, 1] + 0.5) + 0.3 * x[:, 2])
    return mu, 1.7 * mu * (1.0 + 0.2 * np.cos(4.0 * x[:, 1]))


LIVE = {
    "box-hllc-minmod": lambda: cases.periodic_box(7, "HLLC", "Minmod", seed=21),
    "box-roe-vanleer": lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=22),
    "box-ausm-minmod": lambda: cases.periodic_box(6, "AUSMPlusUp", "Minmod", seed=23),
    "box-rusanov-vanleer": lambda: cases.periodic_box(6, "Rusanov", "vanLeer", seed=29),
    "box-rusanov-mrf": lambda: cases.periodic_box(5, "Rusanov", "Minmod", seed=31).with_mrf((10.0, 60.0, 0.0), velocity=(0.0, 0.0, 35.0)),
    "box-hllc-upwind": lambda: cases.periodic_box(5, "HLLC", "upwind", seed=24),
    "box-ragged": lambda: cases.periodic_box(5, "HLLC", "vanLeer", seed=25, nz=3),
    "bump": lambda: cases.bump(15, 10),
    "onera": lambda: cases.onera_box(9),
    "scrambled-hllc": lambda: cases.scrambled_box(6, "HLLC", "vanLeer", seed=41),
    "scrambled-roe": lambda: cases.scrambled_box(5, "ROE", "Minmod", seed=42),
    # laminar viscous residual (residualsUpdate.H:16-43) + LF viscous Jacobian: cyclic + wall + symmetry + mixed patches,
    # a non-orthogonal mesh (bump: totalPressure / directionMixed inlet) and an unstructured numbering
    "box-viscous-roe": lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=51, mu=0.05),
    "box-viscous-hllc": lambda: cases.periodic_box(5, "HLLC", "Minmod", seed=52, mu=0.2, Pr=1.3),
    "bump-viscous": lambda: cases.bump(15, 10, mu=0.02),
    "scrambled-viscous": lambda: cases.scrambled_box(5, "HLLC", "vanLeer", seed=53, mu=0.1),
    # cyclicAMI (non-conformal periodic pair: every face sees two neighbour faces, weights 0.5/0.5 and 0.7/0.3)
    "box-ami-roe": lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=71, ami_shift=0.5),
    "box-ami-hllc-viscous": lambda: cases.periodic_box(5, "HLLC", "Minmod", seed=72, ami_shift=0.3, mu=0.1),
    # MRF: rotating (+ translating) frame, whole mesh or a zone; all three schemes, viscous, non-orthogonal, scrambled
    "box-mrf-hllc": lambda: cases.periodic_box(6, "HLLC", "vanLeer", seed=81).with_mrf((30.0, -50.0, 80.0), (0.3, 0.5, -0.2), (20.0, 5.0, -10.0)),
    "box-mrf-roe-viscous": lambda: cases.periodic_box(5, "ROE", "Minmod", seed=82, mu=0.1).with_mrf((0.0, 0.0, 120.0), (0.5, 0.6, 0.0),
                                                                                                zone=lambda x: x[:, 0] > 0.45),
    "box-mrf-ausm": lambda: cases.periodic_box(5, "AUSMPlusUp", "vanLeer", seed=83).with_mrf((10.0, 60.0, 0.0), velocity=(0.0, 0.0, 35.0)),
    "bump-mrf": lambda: cases.bump(15, 10).with_mrf((0.0, 0.0, 25.0), (1.5, -3.0, 0.0)),
    "scrambled-mrf": lambda: cases.scrambled_box(5, "HLLC", "vanLeer", seed=84).with_mrf((40.0, 0.0, -70.0), (0.2, 0.2, 0.2)),
    # muEff / alphaEff fields from a (stand-in) turbulence model: cyclic, AMI halo, wall patches, non-orthogonal, scrambled; with MRF
    "box-transport-roe": lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=91, mu=0.05).with_transport(TURB),
    "box-transport-ami-hllc": lambda: cases.periodic_box(5, "HLLC", "Minmod", seed=92, mu=0.1, ami_shift=0.3).with_transport(TURB),
    "bump-transport": lambda: cases.bump(15, 10, mu=0.02).with_transport(TURB),
    "scrambled-transport-mrf": lambda: cases.scrambled_box(5, "HLLC", "vanLeer", seed=93, mu=0.1).with_transport(TURB).with_mrf((40.0, 0.0, -70.0), (0.2, 0.2, 0.2)),
    # rotational cyclic pair (90-degree sector): U, gradients and the rhoU operand rotated across the pair
    "rot-hllc-vanleer": lambda: cases.rot_box(6, "HLLC", "vanLeer", seed=62),
    "rot-roe-minmod": lambda: cases.rot_box(5, "ROE", "Minmod", seed=63, nz=4),
    "rot-ausm-mrf": lambda: cases.rot_box(5, "AUSMPlusUp", "vanLeer", seed=64).with_mrf((0.0, 0.0, 60.0)),
    "rot-viscous-hllc": lambda: cases.rot_box(5, "HLLC", "Minmod", seed=66, mu=0.1),
    "rot-ami-roe": lambda: cases.rot_box(6, "ROE", "vanLeer", seed=68, ami_shift=0.5),
    "rot-ami-viscous-full": lambda: FULLV(cases.rot_box(5, "HLLC", "Minmod", seed=69, mu=0.1, ami_shift=0.3)),
    # LaxFriedrichJacobian false: full viscous Jacobian (five fvj::laplacian blocks + wall terms) on cyclic / wall / symmetry /
    # mixed patches, AMI and rotational pairs (values stored on coupled patches), non-orthogonal bump, muEff field, scrambled
    "fullvisc-box-roe": lambda: FULLV(cases.periodic_box(6, "ROE", "vanLeer", seed=101, mu=0.05)),
    "fullvisc-box-hllc-transport": lambda: FULLV(cases.periodic_box(5, "HLLC", "Minmod", seed=102, mu=0.2, Pr=1.3).with_transport(TURB)),
    "fullvisc-ami": lambda: FULLV(cases.periodic_box(5, "HLLC", "Minmod", seed=103, ami_shift=0.3, mu=0.1)),
    "fullvisc-rot": lambda: FULLV(cases.rot_box(5, "ROE", "vanLeer", seed=104, mu=0.1)),
    "fullvisc-bump": lambda: FULLV(cases.bump(15, 10, mu=0.02)),
    "fullvisc-scrambled": lambda: FULLV(cases.scrambled_box(5, "HLLC", "vanLeer", seed=105, mu=0.1)),
    # non-uniform patch entries (inlet profiles): fixedValue / inletOutlet rows, totalPressure p0 and totalTemperature T0 rows
    "box-profiles-roe-viscous": lambda: cases.with_inlet_profiles(cases.periodic_box(6, "ROE", "vanLeer", seed=111, mu=0.05), "ymax"),
    "bump-profiles": lambda: cases.with_inlet_profiles(cases.bump(15, 10), "INLE1", "totalPressure"),
    "shocktube-ausm": lambda: cases.shock_tube(64, "AUSMPlusUp"),
    "shocktube-roe": lambda: cases.shock_tube(50, "ROE"),
}


@pytest.mark.parametrize("name", sorted(LIVE))
def test_against_live_oracle(name, gpu_context):
    case = LIVE[name]()
    ref = run_sequence(case.apply(Oracle()), case)
    out = run_sequence(case.apply(gpu_context()), case)
    compare(out, ref)


def test_bitwise_equality_of_reduction_free_kernels(gpu_context):
    """No tolerance at all where no reduction is involved: fluxes, sources, pseudo time step, all 27 LDU arrays,
    SpMV, LU-SGS and block-Jacobi."""
    case = cases.periodic_box(6, "HLLC", "vanLeer", seed=31)
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    bo, bg = o.boundary_get(), g.boundary_get()       # coupled faces report the patchNeighbourField
    for k in ("rho", "U", "p", "T"):
        assert np.array_equal(bo[k], bg[k]), k
    ami = cases.periodic_box(5, "ROE", "vanLeer", seed=32, ami_shift=0.3)
    bo, bg = ami.apply(Oracle()).boundary_get(), ami.apply(gpu_context()).boundary_get()
    for k in ("rho", "U", "p", "T"):
        assert rel_err(bg[k], bo[k]) <= 1e-14, k          # T and rho are interpolated separately on the two sides
    for a, b in zip(g.calc_flux(), o.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(g.residual(), o.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(g.pseudo_dt()[0], o.pseudo_dt()[0])
    g.assemble(); o.assemble()
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    rng = np.random.default_rng(0)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):
        assert np.array_equal(a, b)
    for pk in ("LUSGS", "Jacobi"):
        for a, b in zip(g.precondition(pk, *x), o.precondition(pk, *x)):
            assert np.array_equal(a, b), pk


def test_mrf_is_bitwise_and_applies_the_coriolis_source_once(gpu_context):
    """MRF terms (icsb200_mrf_set): fluxes, sources, pseudo time step, all LDU arrays and the system sources after one and
    after repeated assembles are bit-identical to the oracle; a zero MRF field reproduces the inertial-frame results."""
    case = cases.periodic_box(6, "ROE", "vanLeer", seed=85).with_mrf((30.0, -50.0, 80.0), (0.3, 0.5, -0.2), (20.0, 5.0, -10.0))
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    for a, b in zip(g.calc_flux(), o.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(g.residual(), o.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(g.pseudo_dt()[0], o.pseudo_dt()[0])
    for rep in range(2):
        g.assemble(); o.assemble()
        for a, b in zip(g.source_get(), o.source_get()):
            assert np.array_equal(a, b), rep
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    plain = cases.periodic_box(6, "ROE", "vanLeer", seed=85)
    zero = cases.periodic_box(6, "ROE", "vanLeer", seed=85).with_mrf()
    a, b = run_sequence(plain.apply(gpu_context()), plain), run_sequence(zero.apply(gpu_context()), zero)
    for k in EXACT_KEYS + SOLVE_KEYS + STATE_KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_transport_fields_bitwise(gpu_context):
    """icsb200_transport_set: viscous sources, all LDU arrays and the matrix product are bit-identical to the oracle with
    variable muEff / alphaEff; a uniform field reproduces the laminar constants; NULL switches back."""
    case = cases.periodic_box(6, "ROE", "vanLeer", seed=95, mu=0.05).with_transport(TURB)
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    for a, b in zip(g.calc_flux(), o.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(g.residual(), o.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(g.pseudo_dt()[0], o.pseudo_dt()[0])
    g.assemble(); o.assemble()
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    lam = cases.periodic_box(6, "ROE", "vanLeer", seed=95, mu=0.05)
    gam = lam.Cp / (lam.Cp - lam.R)
    uni = cases.periodic_box(6, "ROE", "vanLeer", seed=95, mu=0.05).with_transport(lambda x: (np.full(len(x), 0.05), np.full(len(x), gam * (0.05 / lam.Pr))))
    a, b = run_sequence(lam.apply(gpu_context()), lam), run_sequence(uni.apply(gpu_context()), uni)
    for k in EXACT_KEYS + SOLVE_KEYS + STATE_KEYS:
        assert np.array_equal(a[k], b[k]), k
    g.transport_set()   # back to the laminar constants
    g.calc_flux()
    ref = lam.apply(gpu_context()); ref.calc_flux()
    for x, y in zip(g.residual(), ref.residual()):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("name", ["fullvisc-box-roe", "fullvisc-ami", "fullvisc-rot", "fullvisc-bump"])
def test_full_viscous_jacobian_bitwise(name, gpu_context):
    """All 27 LDU arrays and the interface coefficients of the full viscous Jacobian are bit-identical to the oracle."""
    case = LIVE[name]()
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    for api in (o, g):
        api.calc_flux(); api.residual(); api.pseudo_dt(); api.assemble()
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
        assert np.array_equal(g.matrix_get_interfaces(blk), o.matrix_get_interfaces(blk)), blk


def test_transient_dual_time_euler_and_backward(gpu_context):
    for ddt in ("Euler", "backward"):
        case = cases.shock_tube(40, "ROE")
        case.schemes.ddt_scheme = capi.DDT_NAMES[ddt]
        case.schemes.delta_t = 2e-6
        o, g = case.apply(Oracle()), case.apply(gpu_context())
        for step in range(3):
            o.new_time_step(); g.new_time_step()
            for it in range(3):
                ro, rg = o.iterate(case.controls), g.iterate(case.controls)
                assert ro.n_iterations == rg.n_iterations
        so, sg = o.state_get(), g.state_get()
        for k in STATE_KEYS:
            assert rel_err(sg[k], so[k]) <= TOL_STATE, (ddt, k)


def test_solver_only_drop_in_with_host_assembled_matrix(gpu_context):
    """icsb200_matrix_set_ldu: a coupledMatrix assembled on the host (here by the oracle) is solved on the device."""
    case = cases.onera_box(7)
    o = case.apply(Oracle())
    o.calc_flux(); src = o.residual(); o.pseudo_dt(); o.assemble()
    g = case.apply(gpu_context())
    for blk in range(9):
        d, u, l = o.matrix_get_ldu(blk)
        if blk == 1:
            g.matrix_set_ldu(blk, d)        # dSByS(0,1) has only a diagonal (convectiveFluxScheme.C:421)
        else:
            g.matrix_set_ldu(blk, d, u, l)
    g.source_set(*src)
    rng = np.random.default_rng(4)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):
        assert np.array_equal(a, b)
    for a, b in zip(g.precondition("LUSGS", *x), o.precondition("LUSGS", *x)):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", ["cyclic", "ami", "rotational"])
def test_solver_only_drop_in_with_interfaces(name, gpu_context):
    """Solver-only drop-in on a case with coupled patches: the host-assembled LDU arrays AND interfacesUpper coefficients
    (icsb200_matrix_set_ldu + icsb200_matrix_set_interfaces) give the oracle's product, preconditioner and solve; the
    device's own interfaces read back equal the oracle's."""
    case = {"cyclic": lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=33), "ami": lambda: cases.periodic_box(5, "HLLC", "vanLeer", seed=34, ami_shift=0.3),
            "rotational": lambda: cases.rot_box(5, "HLLC", "vanLeer", seed=35)}[name]()
    o = case.apply(Oracle())
    o.calc_flux(); src = o.residual(); o.pseudo_dt(); o.assemble()
    full = case.apply(gpu_context())
    full.calc_flux(); full.residual(); full.pseudo_dt(); full.assemble()
    g = case.apply(gpu_context())
    for blk in range(9):
        d, u, l = o.matrix_get_ldu(blk)
        iu = o.matrix_get_interfaces(blk)
        assert np.array_equal(full.matrix_get_interfaces(blk), iu), blk
        if blk == 1:
            g.matrix_set_ldu(blk, d)
        else:
            g.matrix_set_ldu(blk, d, u, l)
        g.matrix_set_interfaces(blk, iu)
    assert np.abs(o.matrix_get_interfaces(8)).max() > 0
    g.source_set(*src)
    rng = np.random.default_rng(4)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):
        assert np.array_equal(a, b)
    for a, b in zip(g.precondition("LUSGS", *x), o.precondition("LUSGS", *x)):
        assert np.array_equal(a, b)
    (dg, dug, deg), rg = g.solve_delta(case.controls)
    (do, duo, deo), ro = o.solve_delta(case.controls)
    assert rg.n_iterations == ro.n_iterations
    for a, b in ((dg, do), (dug, duo), (deg, deo)):
        assert rel_err(a, b) <= TOL_SOLVE


def test_properties_at_scale(gpu_context):
    """Size-independent properties on a mesh too large for the oracle to be practical in a test (1M cells)."""
    case = cases.onera_box(100)
    g = case.apply(gpu_context())
    m = case.mesh
    phi, phiUp, phiEp = g.calc_flux()
    src = g.residual()
    bnd = np.arange(m.n_internal_faces, m.n_faces)
    # conservation: internal faces cancel in pairs (both sides compute the bit-identical flux)
    assert abs(src[0].sum() + phi[bnd].sum()) < 1e-8 * np.abs(phi).sum()
    assert abs(src[2].sum() + phiEp[bnd].sum()) < 1e-8 * np.abs(phiEp).sum()
    g.pseudo_dt(); g.assemble()
    rng = np.random.default_rng(9)
    N = m.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    y = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    Ax, Ay = g.matrix_mul(*x), g.matrix_mul(*y)
    Axy = g.matrix_mul(*[2.0 * a - 3.0 * b for a, b in zip(x, y)])
    for a, b, c in zip(Ax, Ay, Axy):                                           # linearity of the SpMV
        assert rel_err(c, 2.0 * a - 3.0 * b) < 1e-12
    # LU-SGS is a fixed linear operator: P^-1(2x) = 2 P^-1(x) exactly (scaling by 2 is exact in binary fp)
    px = g.precondition("LUSGS", *x)
    p2x = g.precondition("LUSGS", *[2.0 * a for a in x])
    for a, b in zip(px, p2x):
        assert np.array_equal(2.0 * a, b)
    # row sums: A applied to a constant vector equals diag-dominance terms only -> finite and reproducible run to run
    again = g.matrix_mul(*x)
    for a, b in zip(Ax, again):
        assert np.array_equal(a, b)
    info = g.schedule_info()
    assert info["n_levels_fwd"] == 3 * 100 - 2
    # a few fused iterations stay finite and reduce the residual
    r0 = g.iterate(case.controls)
    for _ in range(3):
        r1 = g.iterate(case.controls)
    st = g.state_get()
    assert np.isfinite(st["rho"]).all() and st["rho"].min() > 0
    assert r1.n_iterations >= 1


def test_edge_cases(gpu_context):
    # two cells, one internal face (smallest mesh with a pseudo time step); errors for wrong call order / selectors
    mesh = mt.structured(1, 2, 1, 1)
    sch = capi.default_schemes(flux_scheme="HLLC", pseudo_co_num=2.0)
    case = cases.Case("one", mesh, 287.0, 1005.0, sch, capi.solver_controls("LUSGS", 2, 2, 0, 1e-12, 1e-3), {}, np.array([1e5, 0.9e5]),
                      np.array([[10.0, 0, 0], [12.0, 0, 0]]), np.array([300.0, 295.0]))
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    ro, rg = o.iterate(case.controls), g.iterate(case.controls)
    assert ro.n_iterations == rg.n_iterations
    assert rel_err(g.state_get()["rho"], o.state_get()["rho"]) < 1e-10
    g2 = gpu_context()
    with pytest.raises(capi.ApiError):
        g2.calc_flux.__self__._call("calc_flux", None, None, None)       # no mesh/state yet
    s = capi.default_schemes()
    s.flux_scheme = 9
    with pytest.raises(capi.ApiError, match="Unknown convectiveFluxScheme"):
        g2.schemes_set(s)
    g3 = case.apply(gpu_context())
    g3.calc_flux(); g3.residual(); g3.pseudo_dt(); g3.assemble()
    bad = capi.solver_controls("LUSGS")
    bad.preconditioner = 5
    with pytest.raises(capi.ApiError, match="Unknown preconditioner"):
        g3.solve_delta(bad)


def test_iterate_host_round_trip(gpu_context):
    """The host-buffer entry point (what an OpenFOAM adapter calls) advances the same state as the resident one."""
    case = cases.onera_box(8)
    a, b = case.apply(gpu_context()), case.apply(gpu_context())
    p, U, T = case.p.copy(), case.U.copy(), case.T.copy()
    for _ in range(3):
        a.iterate(case.controls)
        b.iterate_host(case.controls, p, U, T)
    sa = a.state_get()
    assert rel_err(p, sa["p"]) < 1e-9 and rel_err(U, sa["U"]) < 1e-9 and rel_err(T, sa["T"]) < 1e-9


@pytest.mark.parametrize("mode,make", [("tile", lambda: cases.onera_box(20)), ("tile64", lambda: cases.onera_box(20)),
                                       ("tile64", lambda: cases.onera_box(13)), ("tile64", lambda: cases.bump(24, 20)),
                                       ("tile64", lambda: cases.periodic_box(9, "ROE", "vanLeer", seed=61, mu=0.05)),
                                       ("blk", lambda: cases.onera_box(20)), ("blk", lambda: cases.onera_box(13)),
                                       ("blk", lambda: cases.onera_box(33)), ("blk", lambda: cases.bump(24, 20)),
                                       ("blk", lambda: cases.bump(60, 50)),
                                       ("blk", lambda: cases.periodic_box(9, "ROE", "vanLeer", seed=61, mu=0.05)),
                                       ("blk-tilelevels", lambda: cases.onera_box(20)), ("blk-tilelevels", lambda: cases.bump(60, 50)),
                                       ("blk-depth4", lambda: cases.onera_box(33)), ("blk-depth8", lambda: cases.onera_box(33)),
                                       ("level", lambda: cases.onera_box(13))])
def test_lusgs_tile_mode_is_bit_identical(gpu_context, monkeypatch, mode, make):
    """Every LU-SGS schedule must reproduce the sequential sweeps of the oracle (lusgs.C:220-382) bit for bit:
    ICSB200_LUSGS_MODE=blk (block tiles swept in place by k_lusgs_blk — what `auto`, the default, picks on hex-like meshes),
    level (the level pipeline), tile / tile64 (the older blocked wavefront schedules)."""
    # block tiles: "blk" lets the set-up choose (column mode on these small, chain-bound meshes: a CTA sweeps a column's chunks
    # back to back and hands levels over in shared memory); "-tilelevels" forces the tile-level order with a flag per tile,
    # "-depth4/8" the two tile depths
    if mode.startswith("blk-"):
        if mode == "blk-tilelevels":
            monkeypatch.setenv("ICSB200_LUSGS_COLMODE", "0")
        else:
            monkeypatch.setenv("ICSB200_LUSGS_DEPTH", mode[-1])
        mode = "blk"
    monkeypatch.setenv("ICSB200_LUSGS_MODE", mode)
    case = make()
    g = case.apply(gpu_context())
    info = g.schedule_info()
    if mode == "level":
        assert not info["tile_mode"] and not info["blk"]
    else:
        assert info["tile_mode"] and info["n_tiles"] >= (4 if mode == "blk" else 8)
        assert info["tile_tma"] == (mode == "tile64") and info["blk"] == (mode == "blk")
    monkeypatch.delenv("ICSB200_LUSGS_MODE")
    o = case.apply(Oracle())
    for api in (g, o):
        api.calc_flux(); api.residual(); api.pseudo_dt(); api.assemble()
    rng = np.random.default_rng(3)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    for a, b in zip(g.precondition("LUSGS", *x), o.precondition("LUSGS", *x)):
        assert np.array_equal(a, b)
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):
        assert np.array_equal(a, b)
    for _ in range(2):
        rg, ro = g.iterate(case.controls), o.iterate(case.controls)
        assert rg.n_iterations == ro.n_iterations
    assert rel_err(g.state_get()["rho"], o.state_get()["rho"]) < 1e-10


LOCAL_STEP = os.path.join(cases.tutorial_dir("forwardStep") or "/nonexistent", "constant", "polyMesh")


@pytest.mark.skipif(not os.path.isdir(LOCAL_STEP), reason="forwardStep tutorial not found ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_forward_step_c2_polyhedral_mesh(gpu_context):
    """C2 on the reference's own polyhedral mesh: cells with more than 6 faces, rows with > 3 lower neighbours."""
    case = cases.forward_step(LOCAL_STEP)
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    o.new_time_step(); g.new_time_step()
    for a, b in zip(g.calc_flux(), o.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(g.residual(), o.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(g.pseudo_dt()[0], o.pseudo_dt()[0])
    g.assemble(); o.assemble()
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    rng = np.random.default_rng(5)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    x[1][:, 2] = 0
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):
        assert np.array_equal(a, b)
    for a, b in zip(g.precondition("LUSGS", *x), o.precondition("LUSGS", *x)):
        assert np.array_equal(a, b)
    for it in range(3):
        ro, rg = o.iterate(case.controls), g.iterate(case.controls)
        assert ro.n_iterations == rg.n_iterations
    so, sg = o.state_get(), g.state_get()
    for k in STATE_KEYS:
        assert rel_err(sg[k], so[k]) <= TOL_STATE, k


def test_rotational_cyclic_bitwise_and_refusals(gpu_context):
    """Rotational cyclic pairs on the device (local halo slots filled by k_rot_gather: vector triples rotated by forwardT,
    scalars copied): every reduction-free stage bit for bit against the oracle, incl. the viscous terms; rotational
    cyclicAMI likewise."""
    case = cases.rot_box(6, "HLLC", "vanLeer", seed=61)
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    bo, bg = o.boundary_get(), g.boundary_get()
    for k in ("rho", "U", "p", "T"):
        assert np.array_equal(bo[k], bg[k]), k
    for a, b in zip(g.calc_flux(), o.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(g.residual(), o.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(g.pseudo_dt()[0], o.pseudo_dt()[0])
    g.assemble(); o.assemble()
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    rng = np.random.default_rng(0)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):      # the rhoU part of the operand is rotated across the pair
        assert np.array_equal(a, b)
    for pk in ("LUSGS", "Jacobi"):
        for a, b in zip(g.precondition(pk, *x), o.precondition(pk, *x)):
            assert np.array_equal(a, b), pk
    # viscous terms across the pair: transform(forwardT, .) of grad(U) and tauMC, with a muEff field
    visc = cases.rot_box(5, "ROE", "vanLeer", seed=65)
    visc.mu = 0.1
    visc.with_transport(TURB)
    ov, gv = visc.apply(Oracle()), visc.apply(gpu_context())
    gv.calc_flux(); ov.calc_flux()
    for a, b in zip(gv.residual(), ov.residual()):
        assert np.array_equal(a, b)
    # ... and the device reproduces the full annulus the sector stands for (oracle KAT test_rotational_cyclic_known_answers)
    cs, cf, fcells, scells = cases.sector_and_annulus(6, 0.5)
    gs, gf = cs.apply(gpu_context()), cf.apply(gpu_context())
    gs.calc_flux(); gf.calc_flux()
    for a, b in zip(gs.residual(), gf.residual()):
        assert np.abs(a[scells] - b[fcells]).max() <= 1e-12 * np.abs(a).max()
    # rotational cyclicAMI: interpolate, then rotate; viscous terms interpolate the neighbour cells' tauMC tensors
    ami = cases.rot_box(5, "HLLC", "vanLeer", seed=67, mu=0.1, ami_shift=0.3).with_transport(TURB)
    oa, ga = ami.apply(Oracle()), ami.apply(gpu_context())
    for a, b in zip(ga.calc_flux(), oa.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(ga.residual(), oa.residual()):
        assert np.array_equal(a, b)
    ga.pseudo_dt(); oa.pseudo_dt(); ga.assemble(); oa.assemble()
    for blk in range(9):
        for a, b in zip(ga.matrix_get_ldu(blk), oa.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    for a, b in zip(ga.matrix_mul(*x), oa.matrix_mul(*x)) if ami.mesh.n_cells == N else []:
        assert np.array_equal(a, b)


LOCAL_VKI = os.path.join(cases.tutorial_dir("VKI-LS89") or "/nonexistent", "constant", "polyMesh")


@pytest.mark.skipif(not os.path.isdir(LOCAL_VKI), reason="VKI-LS89 tutorial not found ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_vki_ls89_c5_shipped_mesh(gpu_context):
    """C5 (i) on the reference's own mesh: translational cyclic pair on the device, no-slip isothermal blade, laminar
    viscous residual + LF viscous Jacobian, ROE + vanLeer, GMRES(8)/LU-SGS — every reduction-free stage bit for bit, then
    the residual history and fields of 8 outer iterations."""
    case = cases.vki_ls89(LOCAL_VKI)
    o, g = case.apply(Oracle()), case.apply(gpu_context())
    for a, b in zip(g.calc_flux(), o.calc_flux()):
        assert np.array_equal(a, b)
    for a, b in zip(g.residual(), o.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(g.pseudo_dt()[0], o.pseudo_dt()[0])
    g.assemble(); o.assemble()
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), o.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    rng = np.random.default_rng(6)
    N = case.mesh.n_cells
    x = (rng.standard_normal(N), rng.standard_normal((N, 3)), rng.standard_normal(N))
    x[1][:, 2] = 0
    for a, b in zip(g.matrix_mul(*x), o.matrix_mul(*x)):
        assert np.array_equal(a, b)
    for pk in ("LUSGS", "Jacobi"):
        for a, b in zip(g.precondition(pk, *x), o.precondition(pk, *x)):
            assert np.array_equal(a, b), pk
    for it in range(8):
        ro, rg = o.iterate(case.controls), g.iterate(case.controls)
        assert ro.n_iterations == rg.n_iterations, it
        assert rel_err(list(rg.s_init) + list(rg.v_init)[:2], list(ro.s_init) + list(ro.v_init)[:2]) <= 1e-8, it
    so, sg = o.state_get(), g.state_get()
    for k in STATE_KEYS + ["p", "T"]:
        assert rel_err(sg[k], so[k]) <= TOL_STATE, k
