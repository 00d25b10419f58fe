"""A second, independent reading of the inviscid approximate Jacobian (convectiveFluxScheme::addFluxTerms / addDissipationJacobian /
addTemporalTerms, convectiveFluxScheme.C:366-546, with blockFvMatrix::insertBlock / insertDissipationBlock,
blockFvMatrix.C:211-326) in numpy against all nine LDU sub-blocks of the oracle's coupledMatrix: upper and lower coefficients on
the faces, and the diagonals of the cells, that do not see a boundary.  Same purpose as tests/test_flux_second_reading.py."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle
from tests.test_flux_second_reading import face_states, interior_faces


def second_reading(mesh, s, st, R, Cp, rdt):
    """-> {block: (diag contribution per cell, upper, lower)} with blocks numbered as icsb200_matrix_get_ldu"""
    F, N = mesh.n_internal_faces, mesh.n_cells
    own, nei = mesh.owner[:F], mesh.neighbour
    magSf = mesh.magSf[:F]
    n = mesh.Sf[:F] / magSf[:, None]
    g = s["gamma"]
    V = lambda a: a[:, None]
    T = lambda a: a[:, None, None]
    outer = lambda a, b: a[:, :, None] * b[:, None, :]
    eye = np.eye(3)[None]
    side = {}
    for tag in ("l", "r"):
        U, E = s["U_" + tag][:F], s["E_" + tag][:F]
        theta = 0.5 * (g - 1) * (U * U).sum(1)
        a1, a2 = g * E - theta, g - 1
        proj = (U * n).sum(1)
        side[tag] = {4: n, 6: n * V(theta) - U * V(proj), 8: outer(U, n) - a2 * outer(n, U) + T(proj) * eye, 7: n * a2,
                     2: proj * (theta - a1), 5: n * V(a1) - a2 * U * V(proj), 3: g * proj}
    c = np.sqrt(g * R * st["T"])
    w = mesh.weights[:F]
    lam = (w * c[own] + (1 - w) * c[nei]) + np.abs(((V(w) * st["U"][own] + V(1 - w) * st["U"][nei]) * n).sum(1))
    out = {}
    shape = {0: (), 1: (), 2: (), 3: (), 4: (3,), 5: (3,), 6: (3,), 7: (3,), 8: (3, 3)}
    for b, sh in shape.items():
        ex = (slice(None),) + (None,) * len(sh)
        upper = np.zeros((F,) + sh)
        lower = np.zeros((F,) + sh)
        if b in side["l"]:
            upper += 0.5 * magSf[ex] * side["r"][b]
            lower += -0.5 * magSf[ex] * side["l"][b]
        diag = np.zeros((N,) + sh)
        np.subtract.at(diag, own, lower)                     # negSumDiag of the convective part
        np.subtract.at(diag, nei, upper)
        if b in (0, 3, 8):                                   # insertDissipationBlock (operator-=) and the temporal term
            d = 0.5 * magSf * lam
            dd = d[ex] * (eye if b == 8 else 1.0)
            upper -= dd
            lower -= dd
            np.add.at(diag, own, dd)
            np.add.at(diag, nei, dd)
            diag += (rdt * mesh.V)[ex] * (eye if b == 8 else 1.0)
        out[b] = (diag.reshape(N, -1), upper.reshape(F, -1), lower.reshape(F, -1))
    return out


@pytest.mark.parametrize("make", [lambda: cases.onera_box(7), lambda: cases.periodic_box(7, "ROE", "Minmod", seed=21)])
def test_jacobian_second_reading(make):
    case = make()
    o = case.apply(Oracle())
    o.calc_flux(); o.residual(); rdt, _ = o.pseudo_dt(); o.assemble()
    mesh = case.mesh
    lim = case.schemes.limiter_U
    assert case.schemes.limiter_T == lim
    s = face_states(o, case, lim)
    mine = second_reading(mesh, s, o.state_get(), case.R, case.Cp, rdt)
    f = interior_faces(mesh)
    F = mesh.n_internal_faces
    touches = np.zeros(mesh.n_cells, bool)
    touches[mesh.owner[F:]] = True
    cells = np.flatnonzero(~touches)
    assert len(f) >= 100 and len(cells) >= 100
    for b in range(9):
        d, u, l = o.matrix_get_ldu(b)
        md, mu, ml = mine[b]
        scale = max(np.abs(d).max(), np.abs(u).max(), np.abs(l).max())
        if b == 1:     # dSByS(0,1): no flux term (convectiveFluxScheme.C:424), only addBoundaryTerms touches its diagonal
            assert np.abs(u).max() == 0.0 and np.abs(l).max() == 0.0 and np.abs(d[cells]).max() == 0.0 and np.abs(d).max() > 0.0
            continue
        assert np.abs(u[f] - mu[f]).max() <= 1e-12 * scale, ("upper", b)
        assert np.abs(l[f] - ml[f]).max() <= 1e-12 * scale, ("lower", b)
        assert np.abs(d[cells] - md[cells]).max() <= 1e-12 * scale, ("diag", b)
