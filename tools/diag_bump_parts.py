"""Development aid: LU-SGS schedule and kernel time of every x-slab of an n-way partition of the bump-4M mesh on ONE GPU
(processor patches turned into plain patches: only the schedule and the sweep time matter here)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import cases, capi
from icsfoam_b200.context import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
only = [int(x) for x in sys.argv[2:]]
os.environ["ICSB200_LUSGS_DEBUG"] = "1"
case = cases.bump(1280, 1040)
part, meshes = case.partition(n, "x")
for r, m in enumerate(meshes):
    if only and r not in only:
        continue
    for p in m.patches:
        if p["kind"] == capi.PROCESSOR:
            p["kind"] = capi.PATCH
    g = Context()
    g.mesh_set(m)
    info = g.schedule_info()
    print("rank", r, "cells", m.n_cells, {k: info[k] for k in ("n_levels_fwd", "max_width", "tile_mode", "n_tiles", "n_tile_levels", "blk")}, flush=True)
    g.close()
