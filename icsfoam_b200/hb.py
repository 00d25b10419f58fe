"""Harmonic Balance set-up on the host: the constant operators of HBZone / HBZoneList and the instance-replicated mesh.

Mirrors (names and meaning) src/cfdTools/HB/HBZone.C:46-91 (`setOmegaList`), :190-267 (`calcConditionN`), :270-356
(`updateHBOperators`), HBZoneList.C:32-58 (`findT0`), :253-343 (`setInstants`).  This runs once per case — it is set-up,
not the hot path; the hot path consumes the resulting D through `icsb200_hb_set`.
"""
import numpy as np


def omega_list(frequencies_list, harmonics_list):
    """HBZone::setOmegaList: [0, +k*f ..., reversed(-k*f ...)] (the entries are used as angular frequencies)."""
    if len(frequencies_list) != len(harmonics_list):
        raise ValueError("harmonicsList size must match frequenciesList size")
    pos, neg = [], []
    for f, h in zip(frequencies_list, harmonics_list):
        f = abs(f)
        for k in range(1, int(h) + 1):
            pos.append(k * f)
            neg.append(-k * f)
    return np.array([0.0] + pos + neg[::-1])


def find_T0(omega_lists):
    """HBZoneList::findT0: period of the smallest non-zero |omega| over all zones."""
    m = min(abs(w) for ol in omega_lists for w in ol if abs(w) > 0)
    return 2 * np.pi / m


def condition_number(snapshots, omegas):
    """HBZone::calcConditionN: sqrt(max/min eigenvalue) of Re(E^H E), E[n,k] = exp(i omega_k t_n)."""
    snapshots, omegas = np.asarray(snapshots, float), np.asarray(omegas, float)
    if len(snapshots) < len(omegas):
        raise ValueError("snapshots size must be equal or greater to omegaList size")
    E1 = np.exp(1j * np.outer(snapshots, omegas))
    # E_1.T() is the conjugate transpose for complex matrices (OpenFOAM Matrix::T()); only the real part is kept
    M = (E1.conj().T @ E1).real
    ev = np.linalg.eigvals(M).real
    if ev.min() < 1e-15:
        return 1e15
    return float(np.sqrt(ev.max()) / np.sqrt(ev.min()))


def operators(snapshots, omegas):
    """HBZone::updateHBOperators: EInv[n,k] = exp(i omega_k t_n), E = pinv(EInv), D = -(Im(EInv) A Re(E) + Re(EInv) A Im(E))."""
    snapshots, omegas = np.asarray(snapshots, float), np.asarray(omegas, float)
    EInv = np.exp(1j * np.outer(snapshots, omegas))
    E = np.linalg.pinv(EInv)
    A = np.diag(omegas)
    D = -(EInv.imag @ A @ E.real + EInv.real @ A @ E.imag)
    return EInv, E, np.ascontiguousarray(D)


def phase_lag_operator(snapshots, omegas, ibpa):
    """phaseLagCyclicFvPatchField::phaseLaggedField (phaseLagCyclicFvPatchField.C:282-330): D_pl = Re(EInv M E) with
    M = diag(1, e^{i n IBPA} for the +n harmonics, conjugates for the -n ones): the neighbour value of instance l is
    sum_m D_pl[l][m] * (value of instance m), i.e. the time signal shifted by the inter-blade phase angle."""
    EInv, E, _ = operators(snapshots, omegas)
    nF = len(omegas)
    nH = (nF - 1) // 2
    M = np.zeros((nF, nF), complex)
    M[0, 0] = 1.0
    for n in range(1, nH + 1):
        M[n, n] = np.cos(n * ibpa) + 1j * np.sin(n * ibpa)
        M[nF - n, nF - n] = np.conj(M[n, n])
    return np.ascontiguousarray((EInv @ (M @ E)).real)


def set_instants(omega_lists, n_instants, selected_period=None, oversampling=False):
    """HBZoneList::setInstants: the snapshot times (uniform over `selectedPeriod`, or over the period in [T0, 5 T0]
    with the smallest worst-zone condition number) and every zone's D matrix."""
    for ol in omega_lists:
        if n_instants != len(ol):
            if not oversampling:
                raise ValueError("specified number of instants is not correct")
            if n_instants < len(ol):
                raise ValueError("specified number of instants is lower than frequency vector dimension")
    if selected_period is not None:
        snaps = np.arange(n_instants) * (selected_period / n_instants)
    else:
        T0 = find_T0(omega_lists)
        periods, Tfi = [T0], T0
        while Tfi <= 5 * T0:
            Tfi += 0.001 * T0
            periods.append(Tfi)
        best, snaps = 1e15, np.zeros(n_instants)
        for TF in periods:
            cand = (TF / n_instants) * np.arange(n_instants)
            worst = max([1e-15] + [condition_number(cand, ol) for ol in omega_lists])
            if worst < best:
                best, snaps = worst, cand
    return snaps, [operators(snaps, ol)[2] for ol in omega_lists]


class ReplicatedMesh:
    """n_instants disconnected copies of one mesh, instance-major in cells, faces and patches — the layout
    `icsb200_hb_set` expects ("subTimeLevelK" meshes of dbnsFullyImplicitHBFoam/createMeshes.H:7-95 as one mesh)."""

    def __init__(self, mesh, n):
        N, F, FT = mesh.n_cells, mesh.n_internal_faces, mesh.n_faces
        NB = FT - F
        self.base, self.n_instants = mesh, n
        self.n_cells, self.n_internal_faces, self.n_faces = n * N, n * F, n * FT
        self.solutionD = list(mesh.solutionD)
        rep_int = lambda a: np.concatenate([a[:F]] * n)
        rep_bnd = lambda a: np.concatenate([a[F:]] * n)
        rep = lambda a: np.ascontiguousarray(np.concatenate([rep_int(a), rep_bnd(a)]))
        off_int = np.repeat(np.arange(n, dtype=np.int32) * N, F)
        off_bnd = np.repeat(np.arange(n, dtype=np.int32) * N, NB)
        self.owner = np.ascontiguousarray(np.concatenate([rep_int(mesh.owner) + off_int, rep_bnd(mesh.owner) + off_bnd]).astype(np.int32))
        self.neighbour = np.ascontiguousarray((np.concatenate([mesh.neighbour] * n) + off_int).astype(np.int32))
        for name in ("Sf", "Cf", "magSf", "weights", "deltaCoeffs", "nonOrthDeltaCoeffs"):
            setattr(self, name, rep(getattr(mesh, name)))
        self.C = np.ascontiguousarray(np.concatenate([mesh.C] * n))
        self.V = np.ascontiguousarray(np.concatenate([mesh.V] * n))
        self.patches = []
        np0 = len(mesh.patches)
        for K in range(n):
            for p in mesh.patches:
                q = dict(p)
                q["name"] = f"{p['name']}@{K}"
                q["start"] = n * F + K * NB + (p["start"] - F)
                if q.get("nbr_patch", -1) >= 0:
                    q["nbr_patch"] = q["nbr_patch"] + K * np0   # cyclic / cyclicAMI partner of the same instance ("ami" tables are patch-local)
                self.patches.append(q)
        self.cell_global = None
        self.face_global = None

    def patch_index(self, name):
        return [p["name"] for p in self.patches].index(name)

    def instance_cells(self, K):
        N = self.base.n_cells
        return slice(K * N, (K + 1) * N)


def replicate(mesh, n_instants):
    return ReplicatedMesh(mesh, n_instants)
