"""decomposePar directories as the multi-GPU input: a processorN tree written in decomposePar's layout is
read back rank by rank, the processor-patch geometry is completed from the neighbour directories, and the P-rank oracle
world on those sub-meshes reproduces the single-domain iteration."""
import os

import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from icsfoam_b200.meshtools import foamcase, read_polymesh
from oracle.pyoracle import Oracle, World


def _case_dir(tmp_path, part_fn, n=(6, 5, 4)):
    pts, faces, own, nei, patches = foamcase.box_polymesh(*n, lo=(-1.0, 0.0, 0.0), hi=(2.0, 1.5, 1.0), warp=0.15)
    case_dir = str(tmp_path)
    foamcase.write_polymesh(os.path.join(case_dir, "constant", "polyMesh"), pts, faces, own, nei, patches)
    whole = read_polymesh(os.path.join(case_dir, "constant", "polyMesh"))
    part = part_fn(whole).astype(np.int32)
    foamcase.decompose(case_dir, pts, faces, own, nei, patches, part)
    return case_dir, whole, part


def test_box_polymesh_is_a_valid_mesh(tmp_path):
    case_dir, whole, _ = _case_dir(tmp_path, lambda m: np.zeros(m.n_cells))
    assert whole.n_cells == 120 and whole.n_internal_faces == 5 * 5 * 4 + 6 * 4 * 4 + 6 * 5 * 3
    assert (whole.owner[: whole.n_internal_faces] < whole.neighbour).all()
    assert (whole.V > 0).all() and np.isclose(whole.V.sum(), 3.0 * 1.5 * 1.0, rtol=1e-12)
    # closed cells: the outward area vectors of every cell sum to zero
    s = np.zeros((whole.n_cells, 3))
    np.add.at(s, whole.owner, whole.Sf)
    np.subtract.at(s, whole.neighbour, whole.Sf[: whole.n_internal_faces])
    assert np.abs(s).max() < 1e-14
    assert (whole.nonOrthDeltaCoeffs[: whole.n_internal_faces] != whole.deltaCoeffs[: whole.n_internal_faces]).any()   # warped: non-orthogonal


@pytest.mark.parametrize("n_parts", [2, 4, 8])
def test_read_decomposed_equals_extract_part(tmp_path, n_parts):
    split = {2: lambda m: (m.C[:, 0] > 0.5).astype(int),
             4: lambda m: (m.C[:, 0] > 0.5).astype(int) + 2 * (m.C[:, 1] > 0.7).astype(int),
             # 8 uneven blocks: ranks with up to seven neighbours? no - face neighbours only (<= 3 here), small and large parts
             8: lambda m: (m.C[:, 0] > 0.2).astype(int) + 2 * (m.C[:, 1] > 0.9).astype(int) + 4 * (m.C[:, 2] > 0.3).astype(int)}[n_parts]
    case_dir, whole, part = _case_dir(tmp_path, split)
    assert foamcase.n_processors(case_dir) == n_parts
    cache = {}
    for r in range(n_parts):
        a = foamcase.read_decomposed(case_dir, r, cache)
        b = whole.extract_part(part, r)
        assert (a.n_cells, a.n_internal_faces, a.n_faces) == (b.n_cells, b.n_internal_faces, b.n_faces)
        assert np.array_equal(a.owner, b.owner) and np.array_equal(a.neighbour, b.neighbour)
        assert np.array_equal(a.cell_global, b.cell_global) and np.array_equal(a.face_global, b.face_global)
        assert [(p["name"], p["kind"], p["start"], p["size"], p["nbr_rank"]) for p in a.patches] == \
               [(p["name"], p["kind"], p["start"], p["size"], p["nbr_rank"]) for p in b.patches]
        for name in ("Sf", "Cf", "magSf", "C", "V", "weights", "deltaCoeffs", "nonOrthDeltaCoeffs"):
            assert np.allclose(getattr(a, name), getattr(b, name), rtol=1e-12, atol=1e-14), name
        for p in a.patches:
            if p["kind"] == capi.PROCESSOR:
                f = np.arange(p["start"], p["start"] + p["size"])
                assert ((a.weights[f] > 0) & (a.weights[f] < 1)).all()


def test_world_on_decomposed_directories_matches_single_domain(tmp_path):
    case_dir, whole, part = _case_dir(tmp_path, lambda m: (m.C[:, 0] > 0.5).astype(int) + 2 * (m.C[:, 2] > 0.5).astype(int))
    rng = np.random.default_rng(5)
    N = whole.n_cells
    p = 1e5 * (1 + 0.05 * rng.random(N))
    T = 300.0 * (1 + 0.05 * rng.random(N))
    U = np.column_stack([120.0 + 10 * rng.random(N), 8 * rng.random(N), 5 * rng.random(N)])
    bcs = {name: {"p": ("zeroGradient", ()), "U": ("zeroGradient", ()), "T": ("zeroGradient", ())} for name in ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")}
    case = cases.Case("decomposed", whole, 287.0, 1005.0, capi.default_schemes(flux_scheme="ROE"), capi.solver_controls(), bcs, p, U, T, mu=0.3)
    ctl = capi.solver_controls("Jacobi", n_directions=5, max_iter=30, tolerance=1e-14, rel_tol=1e-9)
    single = case.apply(Oracle())
    part2, meshes = case.decomposed(case_dir)
    assert np.array_equal(part2, part)
    w = World(4)
    w.mesh_set(meshes)
    for o, m in zip(w.ranks, meshes):
        o.thermo_set(case.R, case.Cp, case.mu, case.Pr)
        o.schemes_set(case.schemes)
        for patch, fields in bcs.items():
            for field, (kind, params) in fields.items():
                o.bc_set(patch, {"p": 0, "U": 1, "T": 2}[field], kind, params)
    w.state_set([p[m.cell_global] for m in meshes], [U[m.cell_global] for m in meshes], [T[m.cell_global] for m in meshes])
    # block-Jacobi preconditioning is decomposition independent: same GMRES history and update up to reduction order
    for _ in range(2):
        rw, rs = w.iterate(ctl, 1), single.iterate(ctl)
        assert rw.n_iterations == rs.n_iterations
        assert np.allclose(list(rw.s_init) + list(rw.v_init), list(rs.s_init) + list(rs.v_init), rtol=1e-9)
    st = single.state_get()
    for o, m in zip(w.ranks, meshes):
        sr = o.state_get()
        for k in ("rho", "rhoU", "rhoE"):
            assert np.abs(sr[k] - st[k][m.cell_global]).max() <= 1e-8 * np.abs(st[k]).max(), k


GLOO_DECOMPOSED = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["ICS_ROOT"])
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
import tests.multi_gpu_check as m
from icsfoam_b200 import capi
case, meshes = m.decomposed_case(rank, world)           # rank 0 writes the case and the processor directories, everybody reads them
mine = meshes[rank]
for p in mine.patches:
    if p["kind"] != capi.PROCESSOR:
        continue
    f = np.arange(p["start"], p["start"] + p["size"])
    send = torch.from_numpy(np.ascontiguousarray(np.column_stack([mine.Cf[f], mine.weights[f]])))
    recv = torch.empty_like(send)
    reqs = [dist.isend(send, p["nbr_rank"]), dist.irecv(recv, p["nbr_rank"])]
    [r.wait() for r in reqs]
    assert torch.allclose(send[:, :3], recv[:, :3], atol=1e-12), "processor faces are not matched in order"
    assert torch.allclose(send[:, 3] + recv[:, 3], torch.ones(len(f), dtype=torch.float64), atol=1e-13), "weights of the two sides do not add up to 1"
n = torch.tensor([mine.n_cells], dtype=torch.int64)
dist.all_reduce(n)
assert int(n) == case.mesh.n_cells
print("rank", rank, "ok")
'''


def test_gloo_ranks_share_a_decomposed_case(tmp_path):
    """the file handshake of tests/multi_gpu_check.py's `decomposed` variant (rank 0 writes, barrier, all ranks read) with 2 gloo ranks"""
    import subprocess
    import sys
    from tests.conftest import ROOT
    script = tmp_path / "gloo_decomposed.py"
    script.write_text(GLOO_DECOMPOSED)
    env = dict(os.environ, ICS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29537", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
