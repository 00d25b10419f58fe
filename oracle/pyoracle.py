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
