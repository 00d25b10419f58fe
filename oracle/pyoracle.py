"""ctypes loader of the CPU oracle (oracle/_build/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED: see oracle/oracle.hpp.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from icsfoam_b200 import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make"], cwd=_HERE, check=True, capture_output=True)
        _LIB = C.CDLL(path)
        _LIB.orc_world_create.restype = C.c_void_p
    return _LIB


class Oracle(capi.Api):
    def __init__(self):
        super().__init__(lib(), "orc_", capi.SHARED_SIGNATURES)
        rc = lib().orc_create(C.byref(self.h))
        assert rc == 0


class World:
    """P oracle contexts run SPMD by P threads with halo exchange — the stand-in for an MPI run of the reference."""

    def __init__(self, n):
        self.n = n
        self._w = C.c_void_p(lib().orc_world_create(n))
        self.ranks = [Oracle() for _ in range(n)]
        for r, o in enumerate(self.ranks):
            lib().orc_attach(o.h, self._w, r)

    def mesh_set(self, meshes):
        n = self.n
        self._keep = []
        ctxs = (C.c_void_p * n)(*[o.h for o in self.ranks])
        ints = lambda xs: (C.c_int * n)(*xs)

        def ptrs(arrs):
            self._keep.append(arrs)
            return (C.c_void_p * n)(*[a.ctypes.data for a in arrs])

        patch_arrays = []
        for m, o in zip(meshes, self.ranks):
            o.mesh = m
            pa = (capi.Patch * len(m.patches))()
            for i, p in enumerate(m.patches):
                pa[i].kind, pa[i].start, pa[i].size = p["kind"], p["start"], p["size"]
                pa[i].nbr_rank, pa[i].nbr_patch = p.get("nbr_rank", -1), p.get("nbr_patch", -1)
                for k, v in enumerate([1, 0, 0, 0, 1, 0, 0, 0, 1]):
                    pa[i].forwardT[k] = v
            patch_arrays.append(pa)
        self._keep.append(patch_arrays)
        pp = (C.c_void_p * n)(*[C.addressof(pa) for pa in patch_arrays])
        sd = (C.c_int * 3)(*meshes[0].solutionD)
        rc = lib().orc_world_mesh_set(
            ctxs, n, ints([m.n_cells for m in meshes]), ints([m.n_internal_faces for m in meshes]),
            ints([m.n_faces for m in meshes]), ptrs([m.owner for m in meshes]), ptrs([m.neighbour for m in meshes]),
            ptrs([m.Sf for m in meshes]), ptrs([m.magSf for m in meshes]), ptrs([m.weights for m in meshes]),
            ptrs([m.deltaCoeffs for m in meshes]), ptrs([m.nonOrthDeltaCoeffs for m in meshes]), ptrs([m.C for m in meshes]),
            ptrs([m.V for m in meshes]), ptrs([m.Cf for m in meshes]), ints([len(m.patches) for m in meshes]), pp, sd)
        assert rc == 0, rc

    def state_set(self, ps, Us, Ts):
        n = self.n
        ctxs = (C.c_void_p * n)(*[o.h for o in self.ranks])
        keep = [np.ascontiguousarray(a, np.float64) for a in list(ps) + list(Us) + list(Ts)]
        P = (C.c_void_p * n)(*[a.ctypes.data for a in keep[:n]])
        U = (C.c_void_p * n)(*[a.ctypes.data for a in keep[n:2 * n]])
        T = (C.c_void_p * n)(*[a.ctypes.data for a in keep[2 * n:]])
        lib().orc_world_state_set(ctxs, n, P, U, T)

    def iterate(self, ctl, n_iter=1):
        n = self.n
        ctxs = (C.c_void_p * n)(*[o.h for o in self.ranks])
        res = capi.Residuals()
        rc = lib().orc_world_iterate(ctxs, n, C.byref(ctl), n_iter, C.byref(res))
        assert rc == 0, rc
        return res


def debug_reconstruct(o, lim, cells):
    FT = o.mesh.n_faces
    L, R = np.zeros(FT), np.zeros(FT)
    cells = np.ascontiguousarray(cells, np.float64)
    lib().orc_debug_reconstruct(o.h, int(lim), capi.dptr(cells), capi.dptr(L), capi.dptr(R))
    return L, R


def debug_grad(o, cells):
    g = np.zeros((o.mesh.n_cells, 3))
    cells = np.ascontiguousarray(cells, np.float64)
    lib().orc_debug_grad(o.h, capi.dptr(cells), capi.dptr(g))
    return g


class HB:
    """The reference-structured Harmonic Balance system: nO oracle contexts (one per time-instance mesh) and the global
    (2 nO, nO) coupled solve of dbnsFullyImplicitHBFoam (oracle/oracle_hb.cpp).  Arrays are instance-major."""

    def __init__(self, hbcase, inst=None):
        L = lib()
        L.orc_hb_create.restype = C.c_void_p
        self.case = hbcase
        self.n = hbcase.n_instants
        self.inst = inst if inst is not None else [c.apply(Oracle()) for c in hbcase.instances]
        ctxs = (C.c_void_p * self.n)(*[o.h for o in self.inst])
        self._h = C.c_void_p(L.orc_hb_create(ctxs, self.n))
        N = hbcase.base.mesh.n_cells
        self.N = N
        for z in range(hbcase.D.shape[0]):
            D = np.ascontiguousarray(hbcase.D[z])
            if hbcase.zone_of_cell is None:
                n_cells, cells = -2, None
            else:
                cells = np.ascontiguousarray(np.nonzero(hbcase.zone_of_cell == z)[0], np.int32)
                n_cells = int(cells.size)
            cyl = int(hbcase.cyl_coords[z]) if hbcase.cyl_coords is not None else 0
            ax = np.ascontiguousarray(hbcase.rotation_axis[3 * z:3 * z + 3], np.float64) if cyl else None
            ce = np.ascontiguousarray(hbcase.rotation_centre[3 * z:3 * z + 3], np.float64) if cyl else None
            rc = L.orc_hb_zone_add(self._h, capi.dptr(D), n_cells, capi.iptr(cells), cyl, capi.dptr(ax), capi.dptr(ce))
            assert rc == 0

        for patch, Dpl in getattr(hbcase, "phase_lag_operators", lambda: [])():
            rc = L.orc_hb_phaselag_set(self._h, int(patch), capi.dptr(np.ascontiguousarray(Dpl)))
            assert rc == 0

    def _chk(self, rc, what):
        if rc != 0:
            raise capi.ApiError(f"orc_hb_{what} failed ({rc})")

    def sources(self):
        n, N = self.n, self.N
        a, b, c = np.zeros(n * N), np.zeros((n * N, 3)), np.zeros(n * N)
        self._chk(lib().orc_hb_sources(self._h, capi.dptr(a), capi.dptr(b), capi.dptr(c)), "sources")
        return a, b, c

    def assemble(self):
        self._chk(lib().orc_hb_assemble(self._h), "assemble")

    def residual(self):
        """sources R*V + HB source of every instance after assemble()"""
        n, N = self.n, self.N
        a, b, c = np.zeros(n * N), np.zeros((n * N, 3)), np.zeros(n * N)
        self._chk(lib().orc_hb_system_sources(self._h, capi.dptr(a), capi.dptr(b), capi.dptr(c)), "system_sources")
        return a, b, c

    def matrix_get_ldu(self, block):
        parts = [o.matrix_get_ldu(block) for o in self.inst]
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(3))

    def matrix_mul(self, xRho, xRhoU, xRhoE):
        n, N = self.n, self.N
        a, b, c = np.zeros(n * N), np.zeros((n * N, 3)), np.zeros(n * N)
        self._chk(lib().orc_hb_matrix_mul(self._h, capi.dptr(np.ascontiguousarray(xRho)), capi.dptr(np.ascontiguousarray(xRhoU)),
                                          capi.dptr(np.ascontiguousarray(xRhoE)), capi.dptr(a), capi.dptr(b), capi.dptr(c)), "matrix_mul")
        return a, b, c

    def precondition(self, kind, xRho, xRhoU, xRhoE):
        if isinstance(kind, str):
            kind = capi.PRECOND_NAMES[kind]
        a, b, c = (np.array(xRho, dtype=np.float64, order="C"), np.array(xRhoU, dtype=np.float64, order="C"),
                   np.array(xRhoE, dtype=np.float64, order="C"))
        self._chk(lib().orc_hb_precondition(self._h, kind, capi.dptr(a), capi.dptr(b), capi.dptr(c)), "precondition")
        return a, b, c

    def solve_delta(self, ctl):
        n, N = self.n, self.N
        a, b, c = np.zeros(n * N), np.zeros((n * N, 3)), np.zeros(n * N)
        self._chk(lib().orc_hb_solve_delta(self._h, C.byref(ctl), capi.dptr(a), capi.dptr(b), capi.dptr(c)), "solve_delta")
        return (a, b, c), self.residuals()

    def iterate(self, ctl, n_iter=1):
        self._chk(lib().orc_hb_iterate(self._h, C.byref(ctl), int(n_iter)), "iterate")
        return self.residuals()

    def residuals(self):
        n = self.n
        out = {"s_init": np.zeros(2 * n), "v_init": np.zeros(3 * n), "s_final": np.zeros(2 * n), "v_final": np.zeros(3 * n)}
        it = C.c_int()
        self._chk(lib().orc_hb_residuals_get(self._h, *[capi.dptr(out[k]) for k in ("s_init", "v_init", "s_final", "v_final")],
                                             C.byref(it)), "residuals_get")
        out["n_iterations"] = it.value
        return out

    def state_get(self):
        parts = [o.state_get() for o in self.inst]
        return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}

    def pseudo(self):
        n, N = self.n, self.N
        a, b = np.zeros(n * N), np.zeros(n * N)
        self._chk(lib().orc_hb_pseudo(self._h, capi.dptr(a), capi.dptr(b)), "pseudo")
        return a, b


class HBWorld:
    """P ranks, each holding the Harmonic Balance system of its partition (one HBCase per rank): the stand-in for an MPI run of
    dbnsFullyImplicitHBFoam.  Instance K of every rank forms one World view (halo exchange between the ranks' K-th contexts)."""

    def __init__(self, hbcases):
        self.n = len(hbcases)
        nO = hbcases[0].n_instants
        self._w = C.c_void_p(lib().orc_world_create(self.n))
        ctxs = [[Oracle() for _ in range(nO)] for _ in range(self.n)]
        for r in range(self.n):
            for o in ctxs[r]:
                lib().orc_attach(o.h, self._w, r)
        fid = {"p": capi.FIELD_P, "U": capi.FIELD_U, "T": capi.FIELD_T}
        for K in range(nO):
            view = World.__new__(World)
            view.n, view._w, view.ranks = self.n, self._w, [ctxs[r][K] for r in range(self.n)]
            insts = [hc.instances[K] for hc in hbcases]
            view.mesh_set([c.mesh for c in insts])
            for o, c in zip(view.ranks, insts):
                o.thermo_set(c.R, c.Cp, c.mu, c.Pr)
                o.schemes_set(c.schemes)
                names = [p["name"] for p in c.mesh.patches]
                for patch, fields in c.bcs.items():
                    if patch in names:
                        for field, (kind, params) in fields.items():
                            o.bc_set(patch, fid[field], kind, params)
            view.state_set([c.p for c in insts], [c.U for c in insts], [c.T for c in insts])
        self.ranks = [HB(hc, inst=ctxs[r]) for r, hc in enumerate(hbcases)]

    def iterate(self, ctl, n_iter=1):
        hbs = (C.c_void_p * self.n)(*[h._h for h in self.ranks])
        rc = lib().orc_world_hb_iterate(hbs, self.n, C.byref(ctl), int(n_iter))
        assert rc == 0, rc
        return [h.residuals() for h in self.ranks]
