"""Harmonic Balance: the CUDA path — nO time instances run as ONE replicated mesh behind icsb200_hb_set —
against the reference-structured oracle (nO separate instance contexts + the global (2 nO, nO) system of
dbnsFullyImplicitHBFoam, oracle/oracle_hb.cpp).  GPU only (-m gpu).

Bars as in test_gpu_parity.py: kernels without global reductions must be bit-identical (HB sources, all LDU blocks with
the HB diagonal, the coupled matrix product, LU-SGS with the shared diagonal, the dense 15x15 block-Jacobi); GMRES
increments, residual history and fields after several outer iterations within 1e-8 relative."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import HB
from tests.common import rel_err

pytestmark = pytest.mark.gpu

def _phase_lag(case, ibpa):
    from icsfoam_b200 import hb
    om = 2 * np.pi * 40.0            # hb_box default omega
    harmonics = (case.n_instants - 1) // 2
    return case.with_phase_lag("xmin", "xmax", ibpa, hb.omega_list([om], [harmonics]))


def _rotor(case, viscous=False):
    """HB in the relative frame of a rotor: every time instance sees the same MRF fields (and muEff / alphaEff fields)"""
    for c in case.instances:
        c.with_mrf((0.0, 0.0, 80.0), (0.5, 0.6, 0.0))
        if viscous:
            c.with_transport(lambda x: (0.05 * (1.5 + np.sin(3.0 * x[:, 0])), 0.08 * (1.2 + np.cos(2.0 * x[:, 1]))))
    return case


CASES = {
    "roe-allmesh": lambda: cases.hb_box(6, 3, flux="ROE"),
    "hllc-zoned": lambda: cases.hb_box(5, 3, flux="HLLC", limiter="Minmod", zoned=True, seed=5),
    "roe-cyl": lambda: cases.hb_box(5, 3, flux="ROE", cyl=True, seed=7),
    "ausm-5-instants": lambda: cases.hb_box(4, 5, flux="AUSMPlusUp", seed=9),
    "roe-viscous": lambda: cases.hb_box(5, 3, flux="ROE", seed=11, mu=0.05),   # C5: laminar viscous + HB
    "roe-mrf": lambda: _rotor(cases.hb_box(5, 3, flux="ROE", seed=17)),
    "hllc-mrf-transport-phaselag": lambda: _phase_lag(_rotor(cases.hb_box(4, 3, flux="HLLC", seed=19, mu=0.05), viscous=True), 0.4),
    # phaseLagCyclic pair: the neighbour values of rho p U E H c mix the time instances through D_pl = Re(EInv M(IBPA) E)
    "roe-phaselag": lambda: _phase_lag(cases.hb_box(5, 3, flux="ROE", seed=13), 0.7),
    "hllc-phaselag-viscous": lambda: _phase_lag(cases.hb_box(4, 5, flux="HLLC", limiter="Minmod", seed=15, mu=0.05), -1.1),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_hb_piecewise_bitwise(name, gpu_context):
    case = CASES[name]()
    H = HB(case)
    g = case.apply(gpu_context())
    NT = case.mesh.n_cells
    # sources: flux residual + HB source (outerLoop.H:28-30, residualsUpdate.H:72-74)
    g.calc_flux()
    g.residual()
    rdt_g, co_g = g.pseudo_dt()
    g.assemble()
    H.assemble()
    for a, b in zip(g.source_get(), H.residual()):      # the system sources: R*V + HB source (+ the MRF Coriolis term)
        assert np.array_equal(a, b)
    assert np.array_equal(rdt_g, H.pseudo()[0])
    # all 27 LDU arrays of every instance, with V D[J][J] on the diagonals (HBZone.C:435-518)
    for blk in range(9):
        for a, b in zip(g.matrix_get_ldu(blk), H.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    rng = np.random.default_rng(11)
    x = (rng.standard_normal(NT), rng.standard_normal((NT, 3)), rng.standard_normal(NT))
    # coupled product including the inter-instance diagonal blocks
    for a, b in zip(g.matrix_mul(*x), H.matrix_mul(*x)):
        assert np.array_equal(a, b)
    # LU-SGS with the rDiagCoeff shared by all instances (lusgs.C:50-123)
    for a, b in zip(g.precondition("LUSGS", *x), H.precondition("LUSGS", *x)):
        assert np.array_equal(a, b)
    # dense (5 nO)^2 block-Jacobi
    for a, b in zip(g.precondition("Jacobi", *x), H.precondition("Jacobi", *x)):
        assert rel_err(a, b) <= 1e-12
    # GMRES on the global system
    (dr, dru, dre), res = g.solve_delta(case.controls)
    (orr, oru, ore), ores = H.solve_delta(case.controls)
    assert res.n_iterations == ores["n_iterations"]
    for a, b in ((dr, orr), (dru, oru), (dre, ore)):
        assert rel_err(a, b) <= 1e-8
    gr = g.hb_residuals()
    for k in ("s_init", "v_init", "s_final", "v_final"):
        assert rel_err(gr[k], ores[k]) <= 1e-8, k


@pytest.mark.parametrize("name,precond", [("roe-allmesh", "LUSGS"), ("hllc-zoned", "LUSGS"), ("roe-cyl", "Jacobi"), ("roe-viscous", "LUSGS"),
                                          ("roe-phaselag", "LUSGS"), ("hllc-phaselag-viscous", "LUSGS")])
def test_hb_outer_iterations(name, precond, gpu_context):
    case = CASES[name]()
    ctl = capi.solver_controls(precond, n_directions=5, max_iter=10, tolerance=1e-10, rel_tol=1e-3)
    H = HB(case)
    g = case.apply(gpu_context())
    for it in range(6):
        rg = g.iterate(ctl)
        ro = H.iterate(ctl)
        gr = g.hb_residuals()
        assert rg.n_iterations == ro["n_iterations"], it
        assert rel_err(gr["s_init"], ro["s_init"]) <= 1e-8, it
        assert rel_err(gr["v_init"], ro["v_init"]) <= 1e-8, it
    sg, so = g.state_get(), H.state_get()
    for k in ("rho", "rhoU", "rhoE", "p", "T"):
        assert rel_err(sg[k], so[k]) <= 1e-8, k
    # the SER quirk of the reference's HB loop (Courant number multiplied by 0 on the second iteration) is reproduced
    assert np.array_equal(g.pseudo_dt()[1] > 0, np.ones(case.mesh.n_cells, bool))


def test_hb_set_rejects_non_replicated_mesh(gpu_context):
    case = cases.periodic_box(4)
    g = case.apply(gpu_context())
    D = np.zeros((3, 3))
    with pytest.raises(capi.ApiError):
        g.hb_set(3, D)


def test_hb_off_is_bit_identical_to_plain_path(gpu_context):
    """n_instants = 1 switches HB off; and a replicated mesh WITHOUT hb_set is just nO independent copies."""
    case = cases.hb_box(5, 3)
    g = gpu_context()
    g.mesh_set(case.mesh)
    b = case.base
    g.thermo_set(b.R, b.Cp, b.mu, b.Pr)
    g.schemes_set(case.schemes)
    fid = {"p": capi.FIELD_P, "U": capi.FIELD_U, "T": capi.FIELD_T}
    for K, inst in enumerate(case.instances):
        for patch, fields in inst.bcs.items():
            for field, (kind, params) in fields.items():
                g.bc_set(f"{patch}@{K}", fid[field], kind, params)
    g.state_set(case.p, case.U, case.T)
    g.calc_flux()
    src = g.residual()
    from oracle.pyoracle import Oracle
    N = b.mesh.n_cells
    for K, inst in enumerate(case.instances):
        o = inst.apply(Oracle())
        o.calc_flux()
        for a, r in zip(src, o.residual()):
            assert np.array_equal(a[K * N:(K + 1) * N], r)


class GpuHB:
    """The product driven through the HB-flavoured call names of tests/golden/make_golden_hb.py::run."""

    def __init__(self, g, case):
        self.g, self.case = g, case

    def assemble(self):
        self.g.calc_flux()
        self._src = self.g.residual()
        self._rdt = self.g.pseudo_dt()
        self.g.assemble()

    def residual(self):
        return self._src

    def pseudo(self):
        return self._rdt

    def matrix_get_ldu(self, b):
        return self.g.matrix_get_ldu(b)

    def matrix_mul(self, *x):
        return self.g.matrix_mul(*x)

    def precondition(self, kind, *x):
        return self.g.precondition(kind, *x)

    def iterate(self, ctl):
        r = self.g.iterate(ctl)
        out = self.g.hb_residuals()
        out["n_iterations"] = r.n_iterations
        return out

    def state_get(self):
        return self.g.state_get()


def test_hb_against_golden_fixture(gpu_context):
    import os
    from tests.common import GOLDEN
    from tests.golden import make_golden_hb as mg
    case = mg.make_case()
    got = mg.run(GpuHB(case.apply(gpu_context()), case), case, case.mesh.n_cells)
    gold = np.load(os.path.join(GOLDEN, "hb_box_roe_3instants.npz"))
    for k in ("srcRho", "srcRhoU", "srcRhoE", "rPseudoDeltaT", "diag0", "diag3", "diag8", "y0", "y1", "y2", "z0", "z1", "z2"):
        assert np.array_equal(got[k], gold[k]), k           # no global reduction involved: bit-identical
    assert np.array_equal(got["history"][:, -1], gold["history"][:, -1])
    assert rel_err(got["history"][:, :-1], gold["history"][:, :-1]) <= 1e-8
    for k in ("rho", "rhoU", "rhoE"):
        assert rel_err(got[k], gold[k]) <= 1e-8, k


import os  # noqa: E402

LOCAL_VKI = os.path.join(cases.tutorial_dir("VKI-LS89") or "/nonexistent", "constant", "polyMesh")


@pytest.mark.skipif(not os.path.isdir(LOCAL_VKI), reason="VKI-LS89 tutorial not found ($ICSFOAM_REF, /root/reference or the copy build() stages under cases_local/)")
def test_hb_vki_ls89_c5(gpu_context):
    """C5 (ii): Harmonic Balance with 3 time instances on the shipped VKI-LS89 mesh (84 177 coupled cells, cyclic pair,
    laminar viscous, ROE): HB sources and pseudo time step bit for bit, then 4 outer iterations of the (2 nO, nO) system."""
    case = cases.vki_hb(LOCAL_VKI, 3)
    H = HB(case)
    g = case.apply(gpu_context())
    g.calc_flux()
    src_g = g.residual()
    rdt_g, _ = g.pseudo_dt()
    g.assemble()
    H.assemble()
    for a, b in zip(src_g, H.residual()):
        assert np.array_equal(a, b)
    assert np.array_equal(rdt_g, H.pseudo()[0])
    for blk in (0, 3, 8):
        for a, b in zip(g.matrix_get_ldu(blk), H.matrix_get_ldu(blk)):
            assert np.array_equal(a, b), blk
    for it in range(4):
        rg = g.iterate(case.controls)
        ro = H.iterate(case.controls)
        gr = g.hb_residuals()
        assert rg.n_iterations == ro["n_iterations"], it
        assert rel_err(gr["s_init"], ro["s_init"]) <= 1e-8, it
        assert rel_err(gr["v_init"][np.arange(9) % 3 != 2], ro["v_init"][np.arange(9) % 3 != 2]) <= 1e-8, it
    sg, so = g.state_get(), H.state_get()
    for k in ("rho", "rhoU", "rhoE"):
        assert rel_err(sg[k], so[k]) <= 1e-8, k
