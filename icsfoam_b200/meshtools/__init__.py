"""Host-side mesh inputs for the icsb200 hot path (block-structured generator, polyMesh reader, partitioner).

Thin ctypes wrapper over meshtools.cpp (libicsmesh.so, built by icsfoam_b200.build).  Produces the flat
fvMesh arrays that `icsb200_mesh_set` takes.  Input tooling only — never inside a timed region.
"""
import ctypes as C
import os

import numpy as np

from ..capi import CYCLIC, CYCLICAMI, EMPTY, PATCH, PROCESSOR, SYMMETRYPLANE, WALL  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libicsmesh.so")
        if not os.path.exists(path):
            from .. import build
            build.build_meshtools()
        _LIB = C.CDLL(path)
        _LIB.icsmesh_structured.restype = C.c_void_p
        _LIB.icsmesh_read_polymesh.restype = C.c_void_p
        _LIB.icsmesh_structured_subbox.restype = C.c_void_p
        _LIB.icsmesh_renumber.restype = C.c_void_p
        _LIB.icsmesh_extract_part.restype = C.c_void_p
        _LIB.icsmesh_error.restype = C.c_char_p
    return _LIB


class Mesh:
    """fvMesh arrays in OpenFOAM's native AoS layout."""

    def __init__(self, handle):
        lib = _lib()
        self._h = C.c_void_p(handle)
        err = lib.icsmesh_error(self._h)
        if err:
            raise RuntimeError("meshtools: " + err.decode())
        sz = (C.c_int * 7)()
        lib.icsmesh_sizes(self._h, sz)
        self.n_cells, self.n_internal_faces, self.n_faces, npatch = sz[0], sz[1], sz[2], sz[3]
        self.solutionD = [sz[4], sz[5], sz[6]]
        N, F, FT = self.n_cells, self.n_internal_faces, self.n_faces
        self.owner = np.empty(FT, np.int32)
        self.neighbour = np.empty(F, np.int32)
        self.Sf = np.empty((FT, 3))
        self.Cf = np.empty((FT, 3))
        self.magSf = np.empty(FT)
        self.weights = np.empty(FT)
        self.deltaCoeffs = np.empty(FT)
        self.nonOrthDeltaCoeffs = np.empty(FT)
        self.C = np.empty((N, 3))
        self.V = np.empty(N)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        lib.icsmesh_arrays(self._h, p(self.owner), p(self.neighbour), p(self.Sf), p(self.Cf), p(self.magSf), p(self.weights),
                           p(self.deltaCoeffs), p(self.nonOrthDeltaCoeffs), p(self.C), p(self.V))
        self.patches = []
        for i in range(npatch):
            o = (C.c_int * 5)()
            name = C.create_string_buffer(64)
            lib.icsmesh_patch(self._h, i, o, name)
            self.patches.append({"name": name.value.decode(), "kind": o[0], "start": o[1], "size": o[2], "nbr_rank": o[3],
                                 "nbr_patch": o[4]})
        self.cell_global = None
        self.face_global = None

    def patch_index(self, name):
        return [p["name"] for p in self.patches].index(name)

    def patch_faces(self, name):
        p = self.patches[self.patch_index(name)]
        return np.arange(p["start"], p["start"] + p["size"])

    def set_cyclic(self, a, b):
        """Turn two equally-sized patches into a translational cyclic pair."""
        ia, ib = self.patch_index(a), self.patch_index(b)
        lib = _lib()
        lib.icsmesh_set_patch_kind(self._h, ia, CYCLIC, ib)
        lib.icsmesh_set_patch_kind(self._h, ib, CYCLIC, ia)
        self.patches[ia].update(kind=CYCLIC, nbr_patch=ib)
        self.patches[ib].update(kind=CYCLIC, nbr_patch=ia)
        self._cyclic_geometry(ia, ib)

    def _cyclic_geometry(self, ia, ib):
        """weights / deltaCoeffs / nonOrthDeltaCoeffs of a translational cyclic pair (face i of one patch matches face i of
        the other): cyclicFvPatch::makeWeights (own/neighbour normal distances) and cyclicFvPatch::delta().  Used by the
        synthetic meshes; polyMesh directories get the same from the reader (meshtools.cpp cyclicGeometry)."""
        for pa, pb in ((self.patches[ia], self.patches[ib]), (self.patches[ib], self.patches[ia])):
            fa = np.arange(pa["start"], pa["start"] + pa["size"])
            fb = np.arange(pb["start"], pb["start"] + pb["size"])
            nfa = self.Sf[fa] / self.magSf[fa, None]
            nfb = self.Sf[fb] / self.magSf[fb, None]
            da = np.abs(np.einsum("ij,ij->i", nfa, self.Cf[fa] - self.C[self.owner[fa]]))
            db = np.abs(np.einsum("ij,ij->i", nfb, self.Cf[fb] - self.C[self.owner[fb]]))
            self.weights[fa] = db / (da + db)
            d = (self.Cf[fa] - self.C[self.owner[fa]]) - (self.Cf[fb] - self.C[self.owner[fb]])
            md = np.linalg.norm(d, axis=1)
            self.deltaCoeffs[fa] = 1.0 / md
            self.nonOrthDeltaCoeffs[fa] = 1.0 / np.maximum(np.einsum("ij,ij->i", nfa, d), 0.05 * md)

    def set_cyclic_rotational(self, a, b, T):
        """Turn two equally-sized patches into a ROTATIONAL cyclic pair: face i of `a` coincides with face i of `b` after
        rotating b's frame by T (forwardT of `a`: a vector seen from b's side becomes T.v on a's side; forwardT of `b` is T^T).
        Weights / deltas as cyclicFvPatch::makeWeights and cyclicFvPatch::delta() = patchD - transform(forwardT, nbrPatchD)."""
        T = np.asarray(T, float).reshape(3, 3)
        ia, ib = self.patch_index(a), self.patch_index(b)
        lib = _lib()
        lib.icsmesh_set_patch_kind(self._h, ia, CYCLIC, ib)
        lib.icsmesh_set_patch_kind(self._h, ib, CYCLIC, ia)
        self.patches[ia].update(kind=CYCLIC, nbr_patch=ib, forwardT=[float(v) for v in T.reshape(-1)])
        self.patches[ib].update(kind=CYCLIC, nbr_patch=ia, forwardT=[float(v) for v in T.T.reshape(-1)])
        for pa, pb, R in ((self.patches[ia], self.patches[ib], T), (self.patches[ib], self.patches[ia], T.T)):
            fa = np.arange(pa["start"], pa["start"] + pa["size"])
            fb = np.arange(pb["start"], pb["start"] + pb["size"])
            nfa = self.Sf[fa] / self.magSf[fa, None]
            nfb = self.Sf[fb] / self.magSf[fb, None]
            dA, dB = self.Cf[fa] - self.C[self.owner[fa]], self.Cf[fb] - self.C[self.owner[fb]]
            da = np.abs(np.einsum("ij,ij->i", nfa, dA))
            db = np.abs(np.einsum("ij,ij->i", nfb, dB))
            self.weights[fa] = db / (da + db)
            d = dA - dB @ R.T
            md = np.linalg.norm(d, axis=1)
            self.deltaCoeffs[fa] = 1.0 / md
            self.nonOrthDeltaCoeffs[fa] = 1.0 / np.maximum(np.einsum("ij,ij->i", nfa, d), 0.05 * md)

    def set_cyclic_ami(self, a, b, shift=0.5):
        """Turn two plane patches into a translational cyclicAMI pair whose faces do not match: the neighbour patch is
        shifted by `shift` cells along the first in-plane direction (periodic wrap), so every face overlaps two
        neighbour faces with weights (1 - shift, shift).  shift = 0 gives a one-to-one AMI (== plain cyclic).
        Stand-in for OpenFOAM's AMIInterpolation on the synthetic meshes; fills weights / deltaCoeffs /
        nonOrthDeltaCoeffs of the patch faces as cyclicAMIFvPatch::makeWeights / delta() do."""
        ia, ib = self.patch_index(a), self.patch_index(b)
        tables = {}
        for (i0, i1) in ((ia, ib), (ib, ia)):
            pa, pb = self.patches[i0], self.patches[i1]
            fa = np.arange(pa["start"], pa["start"] + pa["size"])
            fb = np.arange(pb["start"], pb["start"] + pb["size"])
            na = np.abs(self.Sf[fa[0]] / self.magSf[fa[0]])
            axes = [d for d in range(3) if na[d] < 0.5]          # the two in-plane directions
            def grid(faces):
                u = np.unique(np.round(self.Cf[faces][:, axes[0]], 9)); v = np.unique(np.round(self.Cf[faces][:, axes[1]], 9))
                iu = np.searchsorted(u, np.round(self.Cf[faces][:, axes[0]], 9)); iv = np.searchsorted(v, np.round(self.Cf[faces][:, axes[1]], 9))
                return len(u), len(v), iu, iv
            nu, nv, iua, iva = grid(fa)
            nub, nvb, iub, ivb = grid(fb)
            assert (nu, nv) == (nub, nvb) and nu * nv == pa["size"]
            lookup = -np.ones((nu, nv), np.int64)
            lookup[iub, ivb] = np.arange(pb["size"])
            sgn = 1 if i0 == ia else -1                          # the two sides see opposite shifts
            start, face, weight = [0], [], []
            for i in range(pa["size"]):
                if shift == 0:
                    pairs = [(lookup[iua[i], iva[i]], 1.0)]
                else:
                    j0 = lookup[iua[i], iva[i]]
                    j1 = lookup[(iua[i] + sgn) % nu, iva[i]]
                    pairs = [(j0, 1.0 - shift), (j1, shift)]
                for j, w in pairs:
                    face.append(int(j)); weight.append(float(w))
                start.append(len(face))
            tables[i0] = (np.array(start, np.int32), np.array(face, np.int32), np.array(weight, np.float64))
        lib = _lib()
        lib.icsmesh_set_patch_kind(self._h, ia, CYCLICAMI, ib)
        lib.icsmesh_set_patch_kind(self._h, ib, CYCLICAMI, ia)
        self.patches[ia].update(kind=CYCLICAMI, nbr_patch=ib, ami=tables[ia])
        self.patches[ib].update(kind=CYCLICAMI, nbr_patch=ia, ami=tables[ib])
        for (i0, i1) in ((ia, ib), (ib, ia)):
            pa, pb = self.patches[i0], self.patches[i1]
            start, face, weight = pa["ami"]
            fa = np.arange(pa["start"], pa["start"] + pa["size"])
            fb = pb["start"] + face
            nfa = self.Sf[fa] / self.magSf[fa, None]
            da = np.einsum("ij,ij->i", nfa, self.Cf[fa] - self.C[self.owner[fa]])
            nfb = self.Sf[fb] / self.magSf[fb, None]
            dbj = np.einsum("ij,ij->i", nfb, self.Cf[fb] - self.C[self.owner[fb]])
            deltab = self.Cf[fb] - self.C[self.owner[fb]]
            dn = np.zeros(pa["size"]); dvec = np.zeros((pa["size"], 3))
            for i in range(pa["size"]):
                for k in range(start[i], start[i + 1]):
                    dn[i] += weight[k] * dbj[k]
                    dvec[i] += weight[k] * deltab[k]
            self.weights[fa] = dn / (da + dn)
            d = (self.Cf[fa] - self.C[self.owner[fa]]) - dvec
            md = np.linalg.norm(d, axis=1)
            self.deltaCoeffs[fa] = 1.0 / md
            self.nonOrthDeltaCoeffs[fa] = 1.0 / np.maximum(np.einsum("ij,ij->i", nfa, d), 0.05 * md)

    def set_cyclic_ami_rotational(self, a, b, T, shift=0.5):
        """ROTATIONAL cyclicAMI pair: as set_cyclic_ami, but patch `b` coincides with patch `a` only after rotating b's frame by T
        (forwardT of `a`; forwardT of `b` is T^T).  Neighbour faces are located in the rotated frame; weights / deltas as
        cyclicAMIFvPatch::makeWeights and delta() = patchD - transform(forwardT, interpolate(nbrPatchD))."""
        T = np.asarray(T, float).reshape(3, 3)
        ia, ib = self.patch_index(a), self.patch_index(b)
        tables = {}
        for (i0, i1, R) in ((ia, ib, T), (ib, ia, T.T)):
            pa, pb = self.patches[i0], self.patches[i1]
            fa = np.arange(pa["start"], pa["start"] + pa["size"])
            fb = np.arange(pb["start"], pb["start"] + pb["size"])
            na = np.abs(self.Sf[fa[0]] / self.magSf[fa[0]])
            axes = [d for d in range(3) if na[d] < 0.5]
            Xa, Xb = self.Cf[fa], self.Cf[fb] @ R.T          # b's face centres seen from a's side
            def grid(X):
                u = np.unique(np.round(X[:, axes[0]], 9)); v = np.unique(np.round(X[:, axes[1]], 9))
                return len(u), len(v), np.searchsorted(u, np.round(X[:, axes[0]], 9)), np.searchsorted(v, np.round(X[:, axes[1]], 9))
            nu, nv, iua, iva = grid(Xa)
            nub, nvb, iub, ivb = grid(Xb)
            assert (nu, nv) == (nub, nvb) and nu * nv == pa["size"]
            lookup = -np.ones((nu, nv), np.int64)
            lookup[iub, ivb] = np.arange(pb["size"])
            sgn = 1 if i0 == ia else -1
            start, face, weight = [0], [], []
            for i in range(pa["size"]):
                pairs = [(lookup[iua[i], iva[i]], 1.0)] if shift == 0 else [(lookup[iua[i], iva[i]], 1.0 - shift), (lookup[(iua[i] + sgn) % nu, iva[i]], shift)]
                for j, w in pairs:
                    face.append(int(j)); weight.append(float(w))
                start.append(len(face))
            tables[i0] = (np.array(start, np.int32), np.array(face, np.int32), np.array(weight, np.float64), R)
        lib = _lib()
        lib.icsmesh_set_patch_kind(self._h, ia, CYCLICAMI, ib)
        lib.icsmesh_set_patch_kind(self._h, ib, CYCLICAMI, ia)
        for i0, i1 in ((ia, ib), (ib, ia)):
            st, fc, wt, R = tables[i0]
            self.patches[i0].update(kind=CYCLICAMI, nbr_patch=i1, ami=(st, fc, wt), forwardT=[float(v) for v in R.reshape(-1)])
        for (i0, i1) in ((ia, ib), (ib, ia)):
            pa, pb = self.patches[i0], self.patches[i1]
            start, face, weight = pa["ami"]
            R = np.array(pa["forwardT"]).reshape(3, 3)
            fa = np.arange(pa["start"], pa["start"] + pa["size"])
            fb = pb["start"] + face
            nfa = self.Sf[fa] / self.magSf[fa, None]
            da = np.einsum("ij,ij->i", nfa, self.Cf[fa] - self.C[self.owner[fa]])
            nfb = self.Sf[fb] / self.magSf[fb, None]
            dbj = np.einsum("ij,ij->i", nfb, self.Cf[fb] - self.C[self.owner[fb]])
            deltab = self.Cf[fb] - self.C[self.owner[fb]]
            dn = np.zeros(pa["size"]); dvec = np.zeros((pa["size"], 3))
            for i in range(pa["size"]):
                for k in range(start[i], start[i + 1]):
                    dn[i] += weight[k] * dbj[k]
                    dvec[i] += weight[k] * deltab[k]
            self.weights[fa] = dn / (da + dn)
            d = (self.Cf[fa] - self.C[self.owner[fa]]) - dvec @ R.T
            md = np.linalg.norm(d, axis=1)
            self.deltaCoeffs[fa] = 1.0 / md
            self.nonOrthDeltaCoeffs[fa] = 1.0 / np.maximum(np.einsum("ij,ij->i", nfa, d), 0.05 * md)

    def renumber(self, perm):
        """renumberMesh stand-in: new cell id = perm[old id]; faces re-sorted into upper-triangular order."""
        perm = np.ascontiguousarray(perm, np.int32)
        assert sorted(perm.tolist()) == list(range(self.n_cells))
        return Mesh(_lib().icsmesh_renumber(self._h, perm.ctypes.data_as(C.c_void_p)))

    def extract_part(self, part, rank):
        """decomposePar stand-in: sub-mesh of the cells with part[c] == rank (processor patches appended)."""
        part = np.ascontiguousarray(part, np.int32)
        sub = Mesh(_lib().icsmesh_extract_part(self._h, part.ctypes.data_as(C.c_void_p), int(rank)))
        sub.cell_global = np.empty(sub.n_cells, np.int32)
        sub.face_global = np.empty(sub.n_faces, np.int32)
        _lib().icsmesh_part_maps(sub._h, sub.cell_global.ctypes.data_as(C.c_void_p), sub.face_global.ctypes.data_as(C.c_void_p))
        # processor-patch interpolation factors from the parent geometry (processorFvPatch::makeWeights)
        for p in sub.patches:
            if p["kind"] != PROCESSOR:
                continue
            f = np.arange(p["start"], p["start"] + p["size"])
            gf = sub.face_global[f]
            own_is_owner = self.owner[gf] == sub.cell_global[sub.owner[f]]
            nbr = np.where(own_is_owner, self.neighbour[gf], self.owner[gf])
            nf = sub.Sf[f] / sub.magSf[f, None]
            d_own = np.abs(np.einsum("ij,ij->i", nf, sub.Cf[f] - sub.C[sub.owner[f]]))
            d_nbr = np.abs(np.einsum("ij,ij->i", nf, self.C[nbr] - sub.Cf[f]))
            sub.weights[f] = d_nbr / (d_own + d_nbr)
            d = self.C[nbr] - sub.C[sub.owner[f]]
            md = np.linalg.norm(d, axis=1)
            sub.deltaCoeffs[f] = 1.0 / md
            sub.nonOrthDeltaCoeffs[f] = 1.0 / np.maximum(np.einsum("ij,ij->i", nf, d), 0.05 * md)
        # cyclic pairs kept whole on this rank (decomposePar `preservePatches`): same coupled geometry as in the parent mesh
        for i, (p, q0) in enumerate(zip(sub.patches, self.patches)):
            if "forwardT" in q0:
                p["forwardT"] = list(q0["forwardT"])
            if p["kind"] != CYCLIC or p["size"] == 0:
                continue
            q = sub.patches[p["nbr_patch"]]
            fa = np.arange(p["start"], p["start"] + p["size"])
            fb = np.arange(q["start"], q["start"] + q["size"])
            pa0, pb0 = self.patches[i], self.patches[p["nbr_patch"]]
            if p["size"] != q["size"] or not np.array_equal(sub.face_global[fa] - pa0["start"], sub.face_global[fb] - pb0["start"]):
                raise RuntimeError(f"meshtools: the decomposition splits the cyclic pair {p['name']} (use preserve_cyclic)")
            gf = sub.face_global[fa]
            sub.weights[fa], sub.deltaCoeffs[fa], sub.nonOrthDeltaCoeffs[fa] = self.weights[gf], self.deltaCoeffs[gf], self.nonOrthDeltaCoeffs[gf]
        return sub

    def __del__(self):
        try:
            if self._h:
                _lib().icsmesh_free(self._h)
                self._h = None
        except Exception:
            pass


def structured(nb, nxb, ny, nz, kind=0, lo=(0, 0, 0), hi=(1, 1, 1), grad_y=1.0, amp=0.0,
               patch_kinds=(PATCH,) * 6, patch_names=("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")):
    lo_a = (C.c_double * 3)(*lo)
    hi_a = (C.c_double * 3)(*hi)
    kinds = (C.c_int * 6)(*patch_kinds)
    names = (C.c_char_p * 6)(*[n.encode() for n in patch_names])
    h = _lib().icsmesh_structured(nb, nxb, ny, nz, kind, lo_a, hi_a, C.c_double(grad_y), C.c_double(amp), kinds, names)
    return Mesh(h)


def structured_part(n, parts, rank, kind=0, lo=(0, 0, 0), hi=(1, 1, 1), grad_y=1.0, amp=0.0, patch_kinds=(PATCH,) * 6,
                    patch_names=("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")):
    """Partition `rank` of a parts=(px,py,pz) block decomposition of the single-block structured mesh n=(nx,ny,nz),
    built WITHOUT the global mesh: the rank's box plus one ghost layer is generated with the global point mapping and
    then cut with extract_part, which yields processor patches ordered consistently on both sides."""
    n, parts = np.asarray(n), np.asarray(parts)
    r = np.array([rank % parts[0], (rank // parts[0]) % parts[1], rank // (parts[0] * parts[1])])
    edges = [np.linspace(0, n[d], parts[d] + 1).round().astype(int) for d in range(3)]
    lo_i = np.array([edges[d][r[d]] for d in range(3)])
    hi_i = np.array([edges[d][r[d] + 1] for d in range(3)])
    elo, ehi = np.maximum(lo_i - 1, 0), np.minimum(hi_i + 1, n)
    en = ehi - elo
    kinds, names = list(patch_kinds), list(patch_names)
    for d in range(3):                                   # artificial cuts are not real patches
        if elo[d] > 0:
            kinds[2 * d], names[2 * d] = PATCH, f"cut{2 * d}"
        if ehi[d] < n[d]:
            kinds[2 * d + 1], names[2 * d + 1] = PATCH, f"cut{2 * d + 1}"
    ia = lambda v: (C.c_int * 3)(*[int(x) for x in v])
    h = _lib().icsmesh_structured_subbox(ia(en), ia(elo), ia(n), kind, (C.c_double * 3)(*lo), (C.c_double * 3)(*hi),
                                         C.c_double(grad_y), C.c_double(amp), (C.c_int * 6)(*kinds),
                                         (C.c_char_p * 6)(*[s.encode() for s in names]))
    ext = Mesh(h)
    # owner rank of every cell of the extended box
    gi = np.arange(elo[0], ehi[0]); gj = np.arange(elo[1], ehi[1]); gk = np.arange(elo[2], ehi[2])
    ri = np.searchsorted(edges[0], gi, side="right") - 1
    rj = np.searchsorted(edges[1], gj, side="right") - 1
    rk = np.searchsorted(edges[2], gk, side="right") - 1
    part = (ri[None, None, :] + parts[0] * (rj[None, :, None] + parts[1] * rk[:, None, None])).astype(np.int32).reshape(-1)
    sub = ext.extract_part(part, rank)
    # global cell ids of the local cells (lexicographic in the global grid)
    gid = (gi[None, None, :] + n[0] * (gj[None, :, None] + n[1] * gk[:, None, None])).reshape(-1)
    sub.cell_global = gid[sub.cell_global].astype(np.int64)
    return sub


def read_polymesh(directory):
    return Mesh(_lib().icsmesh_read_polymesh(directory.encode()))   # cyclic pairs (neighbourPatch) get their coupled geometry there


# ---- the named configurations of BASELINE.json (SURVEY.md §8d "Configs as concrete synthetic inputs") ----
def shock_tube(n=500):
    """C1: tutorials/shockTube/system/blockMeshDict:17-32 — hex (n 1 1), all six sides 'patch'."""
    return structured(1, n, 1, 1, 0, (-0.5, -0.25, -0.5), (0.5, 0.25, 0.5),
                      patch_kinds=(PATCH,) * 6, patch_names=("side1", "side2", "wallYmin", "wallYmax", "wallZmin", "wallZmax"))


def bump(nxb=66, ny=54):
    """C3: tutorials/circularArcBump/transonic/system/blockMeshDict — 3 blocks of (nxb x ny x 1), grading 3.5 in y."""
    return structured(3, nxb, ny, 1, 1, (-1.5, 0.0, -0.1), (1.5, 1.0, 0.1), grad_y=3.5, amp=0.1,
                      patch_kinds=(PATCH, PATCH, WALL, WALL, EMPTY, EMPTY),
                      patch_names=("INLE1", "PRES2", "WALL4", "WALL3", "defaultFacesA", "defaultFacesB"))


def onera_box(n=48, nz=None, ny=None):
    """C4: synthetic 3-D transonic box (the shipped OneraM6 mesh is incomplete — SURVEY §0.5):
    bump wall on z-min (slip), symmetryPlane on y-min, freestream elsewhere."""
    ny = ny or n
    nz = nz or n
    return structured(1, n, ny, nz, 2, (-1.0, 0.0, 0.0), (2.0, 3.0, 3.0), amp=0.12,
                      patch_kinds=(PATCH, PATCH, SYMMETRYPLANE, PATCH, WALL, PATCH),
                      patch_names=("inlet", "outlet", "symmetry", "lateral", "wing", "top"))
