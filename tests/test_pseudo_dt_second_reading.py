"""A second, independent reading of the local pseudo time step (dbnsFoam/setCoAndDeltaT.H:45-145) in numpy against the oracle:
lambda = interpolate(c) + |interpolate(U) . n| on the faces, the face-wise max of nonOrthDeltaCoeffs * lambda into both cells, the
coupled-patch and the wall (half-weight, cell values) contributions, division by the pseudo-Courant field.  Same purpose as
tests/test_flux_second_reading.py."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle


def local_pseudo_dt(mesh, R, Cp, st, co, mrf=None):
    F, N = mesh.n_internal_faces, mesh.n_cells
    gamma = Cp / (Cp - R)
    c = np.sqrt(gamma / (1.0 / (R * st["T"])))                   # sqrt(gamma / psi)
    U = st["U"]
    n = mesh.Sf / mesh.magSf[:, None]
    own, nei, w = mesh.owner, mesh.neighbour, mesh.weights
    lin = lambda f, P, Nb, wf: wf * P + (1.0 - wf) * Nb if P.ndim == 1 else wf[:, None] * P + (1.0 - wf)[:, None] * Nb
    mrf = np.zeros(mesh.n_faces) if mrf is None else mrf
    lam = lin(None, c[own[:F]], c[nei], w[:F]) + np.abs((lin(None, U[own[:F]], U[nei], w[:F]) * n[:F]).sum(1) - mrf[:F])
    frdt = mesh.nonOrthDeltaCoeffs[:F] * lam
    rdt = np.zeros(N)
    np.maximum.at(rdt, own[:F], frdt)
    np.maximum.at(rdt, nei, frdt)
    for p in mesh.patches:
        f = np.arange(p["start"], p["start"] + p["size"])
        if p["size"] == 0:
            continue
        fc = own[f]
        if p["kind"] == capi.CYCLIC:
            q = mesh.patches[p["nbr_patch"]]
            nb = own[np.arange(q["start"], q["start"] + q["size"])]             # patchNeighbourField: the cells across the pair
            lam_b = lin(None, c[fc], c[nb], w[f]) + np.abs((lin(None, U[fc], U[nb], w[f]) * n[f]).sum(1) - mrf[f])
            np.maximum.at(rdt, fc, mesh.nonOrthDeltaCoeffs[f] * lam_b)
        elif p["kind"] == capi.WALL:
            np.maximum.at(rdt, fc, 0.5 * mesh.nonOrthDeltaCoeffs[f] * (c[fc] + np.abs((U[fc] * n[f]).sum(1) - mrf[f])))
    return rdt / co


@pytest.mark.parametrize("make", [lambda: cases.onera_box(7), lambda: cases.bump(12, 9), lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=3),
                                  lambda: cases.scrambled_box(5, "HLLC", "vanLeer", seed=4)])
def test_local_pseudo_time_step_second_reading(make):
    case = make()
    o = case.apply(Oracle())
    o.calc_flux()
    o.residual()
    rdt, co = o.pseudo_dt()
    assert np.all(co == case.schemes.pseudo_co_num)              # first iteration: no SER update yet
    mine = local_pseudo_dt(case.mesh, case.R, case.Cp, o.state_get(), co)
    assert np.allclose(rdt, mine, rtol=1e-13, atol=0.0)


def test_local_pseudo_time_step_in_a_rotating_frame():
    for make in (lambda: cases.periodic_box(6, "ROE", "vanLeer", seed=3), lambda: cases.bump(12, 9)):
        case = make().with_mrf((0.0, 0.0, 120.0), (0.5, 0.6, 0.0), (20.0, 5.0, -10.0))
        o = case.apply(Oracle())
        o.calc_flux()
        o.residual()
        rdt, co = o.pseudo_dt()
        fv = case.mrf_fields(case.mesh)[0]
        mine = local_pseudo_dt(case.mesh, case.R, case.Cp, o.state_get(), co, fv)
        plain = local_pseudo_dt(case.mesh, case.R, case.Cp, o.state_get(), co)
        assert np.allclose(rdt, mine, rtol=1e-13, atol=0.0) and not np.allclose(rdt, plain, rtol=1e-6)


def test_switched_evolution_relaxation_second_reading():
    """setCoAndDeltaT.H:1-35: from the third outer iteration on, Co *= clip(|res_{k-2}| / |res_{k-1}|, minDecr, maxIncr), clipped to
    [CoMin, CoMax], with |res| the 2-norm of the (rho, rhoE, |rhoU|) initial residuals of the linear solves."""
    case = cases.onera_box(6)
    sc = case.schemes
    sc.pseudo_co_num, sc.pseudo_co_num_max, sc.pseudo_co_num_min = 5.0, 12.0, 0.5
    o = case.apply(Oracle())
    norms, seen = [], []
    for k in range(7):
        o.calc_flux(); o.residual(); _, co = o.pseudo_dt(); o.assemble()
        _, res = o.solve_delta(case.controls)
        o.update_fields()
        assert co.min() == co.max()
        seen.append(co.max())
        norms.append(np.sqrt(res.s_init[0] ** 2 + res.s_init[1] ** 2 + sum(v * v for v in res.v_init)))
    co, want = 5.0, []
    for k in range(7):
        if k >= 2:
            ratio = max(min(norms[k - 2] / norms[k - 1], sc.pseudo_co_num_max_incr), sc.pseudo_co_num_min_decr)
            co = max(min(co * ratio, sc.pseudo_co_num_max), sc.pseudo_co_num_min)
        want.append(co)
    assert np.allclose(seen, want, rtol=1e-13) and seen[2] > seen[1] and seen[-1] == 12.0


def test_global_pseudo_time_step_second_reading():
    """localTimestepping false: rPseudoDeltaT = max(deltaCoeffs * lambda) / pseudoCoNum over all faces (setCoAndDeltaT.H:147-151); on
    a box whose patches are all zeroGradient the boundary values of c and U are the cell values."""
    from icsfoam_b200 import meshtools as mt
    mesh = mt.structured(1, 6, 5, 4, 0, (0, 0, 0), (1.2, 1.0, 0.8), patch_kinds=(capi.PATCH,) * 6)
    rng = np.random.default_rng(12)
    N = mesh.n_cells
    p, T = 1e5 * (1 + 0.1 * rng.random(N)), 300.0 * (1 + 0.1 * rng.random(N))
    U = np.column_stack([150.0 + 40 * rng.random(N), 30 * rng.standard_normal(N), 30 * rng.standard_normal(N)])
    bcs = {q["name"]: {"p": ("zeroGradient", ()), "U": ("zeroGradient", ()), "T": ("zeroGradient", ())} for q in mesh.patches}
    sch = capi.default_schemes(flux_scheme="HLLC", local_timestepping=0, pseudo_co_num=3.0)
    case = cases.Case("zg", mesh, 287.0, 1005.0, sch, capi.solver_controls(), bcs, p, U, T)
    o = case.apply(Oracle())
    o.calc_flux(); o.residual()
    rdt, _ = o.pseudo_dt()
    st = o.state_get()
    F = mesh.n_internal_faces
    g = case.Cp / (case.Cp - case.R)
    c = np.sqrt(g * case.R * st["T"])
    n = mesh.Sf / mesh.magSf[:, None]
    own, nei, w = mesh.owner, mesh.neighbour, mesh.weights
    lam_i = (w[:F] * c[own[:F]] + (1 - w[:F]) * c[nei]) + np.abs(((w[:F, None] * st["U"][own[:F]] + (1 - w[:F, None]) * st["U"][nei]) * n[:F]).sum(1))
    lam_b = c[own[F:]] + np.abs((st["U"][own[F:]] * n[F:]).sum(1))
    want = max((mesh.deltaCoeffs[:F] * lam_i).max(), (mesh.deltaCoeffs[F:] * lam_b).max()) / 3.0
    assert rdt.min() == rdt.max() and np.isclose(rdt[0], want, rtol=1e-13)


def test_local_time_step_bounding_as_coded():
    """boundLocalTimeStep.H:1-98 runs right after solveForIncr and BEFORE updateFields.H adds the increments (outerLoop.H:88-99,
    dbnsFoam.C:112-118), so `rho` still equals scalarVarsPrevIter[0]: the tests rho < lowerBound rhoPrev and e < lowerBound ePrev
    cannot fire for a positive state and only a non-positive internal energy of the CURRENT state reduces the pseudo-Courant
    number.  The oracle (and the device) follow the reference as coded: cells whose density or energy drops by more than 5 % in
    the update keep their Courant number."""
    case = cases.periodic_box(6, "ROE", "vanLeer", seed=3)
    case.schemes.pseudo_co_num = 0.5
    case.schemes.pseudo_co_num_min = 0.01
    o = case.apply(Oracle())
    prev = o.state_get()
    o.iterate(case.controls)
    new = o.state_get()
    o.calc_flux()
    o.residual()
    _, co1 = o.pseudo_dt()                                    # second outer iteration: no SER update yet
    lb = case.schemes.local_timestepping_lower_bound
    e_of = lambda s: s["rhoE"] / s["rho"] - 0.5 * ((s["rhoU"] / s["rho"][:, None]) ** 2).sum(1)
    dropped = (new["rho"] < lb * prev["rho"]) | (e_of(new) < lb * e_of(prev))
    assert dropped.sum() >= 10 and (e_of(prev) > 0).all() and (e_of(new) > 0).all()
    assert (co1 == 0.5).all()
