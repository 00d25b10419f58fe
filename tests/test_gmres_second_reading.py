"""A second, independent reading of gmres::solveDelta (gmres.C:44-69,772-1110) and the block preconditioners in dense numpy,
against the C++ oracle, restart by restart: the Krylov recurrence with modified Gram-Schmidt, the reference's Givens
convention, the back substitution with stabilise(), the solution update, and the residual bookkeeping (normalisation factors
from A (W - avg W) and the source, component-wise vector residuals) that the reported residual history consists of.
Same purpose as tests/test_flux_second_reading.py: a transcription slip in either reading shows up here."""
import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle
from tests.test_oracle_kat import dense_from_ldu

VSMALL = 1e-300


def givens(h, beta):
    if beta == 0:
        return 1.0, 0.0
    if abs(beta) > abs(h):
        tau = -h / beta
        s = 1.0 / np.sqrt(1.0 + tau * tau)
        return s * tau, s
    tau = -beta / h
    c = 1.0 / np.sqrt(1.0 + tau * tau)
    return c, c * tau


def gmres_solve_delta(A, precond, W, src, n_dirs, n_restarts):
    """W, src: (N, 5) in the order (rho, rhoU, rhoE).  Returns dW and the residual records after every restart."""
    N = W.shape[0]
    mul = lambda x: (A @ x.reshape(-1)).reshape(N, 5)
    dW = W - W.mean(axis=0)                                     # gAverage per scalar field / per vector field (component-wise)
    tmp = mul(dW)
    sN = [np.abs(tmp[:, k]).sum() + np.abs(src[:, k]).sum() + VSMALL for k in (0, 4)]
    vN = (np.linalg.norm(tmp[:, 1:4], axis=1) + np.linalg.norm(src[:, 1:4], axis=1)).sum() + VSMALL
    init = [np.abs(src[:, 0]).sum() / sN[0], np.abs(src[:, 4]).sum() / sN[1]] + list(np.abs(src[:, 1:4]).sum(axis=0) / vN)
    dW = np.zeros_like(W)
    tmp = src.copy()
    history = []
    for _ in range(n_restarts):
        tmp = precond(tmp)
        beta = np.sqrt((tmp * tmp).sum())
        H = np.zeros((n_dirs, n_dirs))
        bh = np.zeros(n_dirs + 1)
        bh[0] = beta
        c, s = np.zeros(n_dirs), np.zeros(n_dirs)
        Vs = []
        for i in range(n_dirs):
            Vs.append(tmp / beta)
            tmp = precond(mul(Vs[i]))
            for j in range(i + 1):
                H[j, i] = (tmp * Vs[j]).sum()
                tmp = tmp - H[j, i] * Vs[j]
            beta = np.sqrt((tmp * tmp).sum())
            for j in range(i):
                Hji = H[j, i]
                H[j, i] = c[j] * Hji - s[j] * H[j + 1, i]
                H[j + 1, i] = s[j] * Hji + c[j] * H[j + 1, i]
            c[i], s[i] = givens(H[i, i], beta)
            bhi = bh[i]
            bh[i] = c[i] * bhi - s[i] * bh[i + 1]
            bh[i + 1] = s[i] * bhi + c[i] * bh[i + 1]
            H[i, i] = c[i] * H[i, i] - s[i] * beta
        yh = np.zeros(n_dirs)
        for i in range(n_dirs - 1, -1, -1):
            acc = bh[i] - sum(H[i, j] * yh[j] for j in range(i + 1, n_dirs))
            d = H[i, i]
            yh[i] = acc / (d - VSMALL if d < 0 else d + VSMALL)          # stabilise(H_ii, VSMALL)
        for i in range(n_dirs):
            dW = dW + yh[i] * Vs[i]
        tmp = src - mul(dW)
        final = [np.abs(tmp[:, 0]).sum() / sN[0], np.abs(tmp[:, 4]).sum() / sN[1]] + list(np.abs(tmp[:, 1:4]).sum(axis=0) / vN)
        history.append((dW.copy(), final))
    return init, history


def dense_preconditioners(A, N):
    blocks = A.reshape(N, 5, N, 5)
    inv_diag = np.stack([np.linalg.inv(blocks[i, :, i, :]) for i in range(N)])
    jacobi = lambda x: np.einsum("nij,nj->ni", inv_diag, x)                      # Jacobi.C:55-132
    # lusgs.C:50-382: (D + L) D^-1 (D + U) with the scalar D = max |diagonal coefficient| of the cell, sweeps in cell order
    rD = 1.0 / np.array([np.abs(np.diag(blocks[i, :, i, :])).max() for i in range(N)])

    def lusgs(x):
        y = np.zeros_like(x)
        for i in range(N):                                                       # forward: y_i = rD_i (x_i - sum_{j<i} A_ij y_j)
            acc = x[i].copy()
            for j in np.flatnonzero(np.abs(blocks[i, :, :i, :]).sum(axis=(0, 2))):
                acc -= blocks[i, :, j, :] @ y[j]
            y[i] = rD[i] * acc
        for i in range(N - 1, -1, -1):                                           # reverse: y_i -= rD_i sum_{j>i} A_ij y_j
            acc = np.zeros(5)
            for j in i + 1 + np.flatnonzero(np.abs(blocks[i, :, i + 1:, :]).sum(axis=(0, 2))):
                acc += blocks[i, :, j, :] @ y[j]
            y[i] = y[i] - rD[i] * acc
        return y

    return {"Jacobi": jacobi, "LUSGS": lusgs}


@pytest.mark.parametrize("precond", ["Jacobi", "LUSGS"])
def test_gmres_second_reading_agrees_restart_by_restart(precond):
    case = cases.onera_box(5)
    o = case.apply(Oracle())
    o.calc_flux(); src = o.residual(); o.pseudo_dt(); o.assemble()
    N = case.mesh.n_cells
    A = dense_from_ldu(o, case.mesh)
    st = o.state_get()
    W = np.column_stack([st["rho"], st["rhoU"], st["rhoE"]])
    b = np.column_stack(src)
    P = dense_preconditioners(A, N)[precond]
    m = 4
    init, history = gmres_solve_delta(A, P, W, b, m, 3)
    for k, (dW_ref, final) in enumerate(history, start=1):
        ctl = capi.solver_controls(precond, n_directions=m, max_iter=k, min_iter=k, tolerance=1e-300, rel_tol=0.0)
        (dr, dru, dre), res = o.solve_delta(ctl)
        assert res.n_iterations == k
        dW = np.column_stack([dr, dru, dre])
        assert np.abs(dW - dW_ref).max() <= 1e-9 * np.abs(dW_ref).max(), (precond, k)
        got_init = list(res.s_init) + list(res.v_init)
        got_final = list(res.s_final) + list(res.v_final)
        assert np.allclose(got_init[:5], init, rtol=1e-10), (precond, k)
        assert np.allclose(got_final[:5], final, rtol=1e-7, atol=1e-14), (precond, k)
    assert history[-1][1][0] < init[0]


def test_smooth_solver_second_reading():
    """smoothSolverCoupled::solveDelta (smoothSolverCoupled.C:385-515) with JacobiSmoother::smooth (JacobiSmoother.C:120-203) in dense
    numpy: nSweeps block-Jacobi sweeps x <- D^-1 (b - (A - D) x) between two residual evaluations, nIterations counted in sweeps,
    the same residual normalisation as GMRES."""
    case = cases.onera_box(5)
    o = case.apply(Oracle())
    o.calc_flux(); src = o.residual(); o.pseudo_dt(); o.assemble()
    N = case.mesh.n_cells
    A = dense_from_ldu(o, case.mesh)
    blocks = A.reshape(N, 5, N, 5)
    D = np.zeros_like(A)
    for i in range(N):
        D[5 * i:5 * i + 5, 5 * i:5 * i + 5] = blocks[i, :, i, :]
    Dinv = np.linalg.inv(D)
    st = o.state_get()
    W = np.column_stack([st["rho"], st["rhoU"], st["rhoE"]])
    b = np.column_stack(src).reshape(-1)
    tmp = (A @ (W - W.mean(axis=0)).reshape(-1)).reshape(N, 5)
    bb = b.reshape(N, 5)
    sN = [np.abs(tmp[:, k]).sum() + np.abs(bb[:, k]).sum() + VSMALL for k in (0, 4)]
    vN = (np.linalg.norm(tmp[:, 1:4], axis=1) + np.linalg.norm(bb[:, 1:4], axis=1)).sum() + VSMALL
    n_sweeps = 2
    x = np.zeros(5 * N)
    for outer in range(1, 4):
        for _ in range(n_sweeps):
            x = Dinv @ (b - (A - D) @ x)
        r = (b - A @ x).reshape(N, 5)
        final = [np.abs(r[:, 0]).sum() / sN[0], np.abs(r[:, 4]).sum() / sN[1]] + list(np.abs(r[:, 1:4]).sum(axis=0) / vN)
        ctl = capi.solver_controls(solver="smoothSolverCoupled", n_sweeps=n_sweeps, max_iter=outer * n_sweeps, min_iter=outer * n_sweeps,
                                   tolerance=1e-300, rel_tol=0.0)
        (dr, dru, dre), res = o.solve_delta(ctl)
        assert res.n_iterations == outer * n_sweeps
        got = np.column_stack([dr, dru, dre]).reshape(-1)
        assert np.abs(got - x).max() <= 1e-11 * np.abs(x).max(), outer
        assert np.allclose(list(res.s_final) + list(res.v_final), final, rtol=1e-9), outer
    assert final[0] < np.abs(bb[:, 0]).sum() / sN[0]                      # the sweeps reduce the residual on this diagonally dominant system
    # one zero-start sweep is the Jacobi preconditioner (Jacobi.C:55-132)
    z = o.precondition("Jacobi", bb[:, 0].copy(), bb[:, 1:4].copy(), bb[:, 4].copy())
    assert np.allclose(np.column_stack(z).reshape(-1), Dinv @ b, rtol=1e-9, atol=1e-12 * np.abs(Dinv @ b).max())
