"""The CUDA product behind the icsb200 C-ABI (libicsb200.so, sm_100a).  No CPU fallback: constructing a
`Context` without the built library or without a B200-class GPU raises."""
import ctypes as C
import os

import numpy as np

from . import capi

_PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None
LIB_PATH = os.path.join(_PKG, "libicsb200.so")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the CUDA extension is the product; there is no fallback path)")
        _LIB = C.CDLL(LIB_PATH)
    return _LIB


def exported_symbols():
    """Names declared in include/icsb200.h (used by the CPU-side 'library loads and exports' test)."""
    import re
    hdr = open(os.path.join(os.path.dirname(_PKG), "include", "icsb200.h")).read()
    return sorted(set(re.findall(r"\b(icsb200_[a-z0-9_]+)\s*\(", hdr)))


class Context(capi.Api):
    def __init__(self, device=0, nccl_id=None, rank=0, n_ranks=1):
        sigs = dict(capi.SHARED_SIGNATURES)
        sigs.update(capi.PRODUCT_SIGNATURES)
        super().__init__(lib(), "icsb200_", sigs)
        idbuf = None
        if nccl_id is not None:
            idbuf = C.create_string_buffer(bytes(nccl_id), 128)
        rc = self._fn["create"](C.byref(self.h), device, idbuf, rank, n_ranks)
        if rc != 0 or not self.h:
            raise RuntimeError(f"icsb200_create failed ({rc}): needs an sm_100 (B200) GPU; there is no CPU fallback")

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        f = lib().icsb200_nccl_unique_id
        f.restype = C.c_int
        f.argtypes = [C.c_void_p]
        rc = f(buf)
        if rc != 0:
            raise RuntimeError("icsb200_nccl_unique_id failed")
        return buf.raw

    def iterate_host(self, ctl, p, U, T):
        res = capi.Residuals()
        self._call("iterate_host", C.byref(ctl), capi.dptr(p), capi.dptr(U), capi.dptr(T), C.byref(res))
        return res

    def launch_count(self):
        return int(self._fn["launch_count"](self.h))

    def timers_reset(self, enable=True):
        self._call("timers_reset", 1 if enable else 0)

    def timers_get(self):
        names = C.create_string_buffer(1024)
        ms = (C.c_double * 32)()
        calls = (C.c_longlong * 32)()
        n = self._fn["timers_get"](self.h, names, 1024, ms, calls, 32)
        nm = names.raw.split(b"\0")[:n]
        return {nm[i].decode(): (ms[i], int(calls[i])) for i in range(n)}

    def timer_begin(self):
        self._call("timer_begin")

    def timer_end(self):
        ms = C.c_double()
        self._call("timer_end", C.byref(ms))
        return ms.value

    def schedule_info(self):
        out = (C.c_int * 8)()
        self._call("schedule_info", out)
        return {"n_levels_fwd": out[0], "n_levels_rev": out[1], "max_width": out[2], "n_positions": out[3],
                "tile_mode": bool(out[4]), "n_tiles": out[5], "n_tile_levels": out[6], "tile_tma": out[7] == 1, "blk": out[7] == 2}

    # ---- Harmonic Balance (icsb200_hb_set): the mesh must be an `hb.ReplicatedMesh`
    def hb_set(self, n_instants, D, zone_of_cell=None, cyl_coords=None, rotation_axis=None, rotation_centre=None):
        D = np.ascontiguousarray(np.asarray(D, np.float64).reshape(-1, n_instants, n_instants))
        nz = D.shape[0]
        zoc = None if zone_of_cell is None else np.ascontiguousarray(zone_of_cell, np.int32)
        cyl = None if cyl_coords is None else np.ascontiguousarray(cyl_coords, np.int32)
        ax = None if rotation_axis is None else np.ascontiguousarray(rotation_axis, np.float64)
        ce = None if rotation_centre is None else np.ascontiguousarray(rotation_centre, np.float64)
        self._call("hb_set", int(n_instants), int(nz), capi.dptr(D), capi.iptr(zoc), capi.iptr(cyl), capi.dptr(ax), capi.dptr(ce))
        self.n_instants = int(n_instants)

    def phaselag_set(self, patch, weights):
        """phaseLagCyclic: row of D_pl for replicated patch `patch` (icsb200_phaselag_set; call before mesh_set)"""
        w = np.ascontiguousarray(weights, np.float64)
        self._call("phaselag_set", int(patch), int(w.size), capi.dptr(w))

    def hb_residuals(self):
        n = self.n_instants
        out = {"s_init": np.zeros(2 * n), "v_init": np.zeros(3 * n), "s_final": np.zeros(2 * n), "v_final": np.zeros(3 * n)}
        self._call("hb_residuals_get", *[capi.dptr(out[k]) for k in ("s_init", "v_init", "s_final", "v_final")])
        return out
