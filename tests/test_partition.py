"""N>1 host logic on CPU: the decomposePar stand-in produces consistent processor patches, the P-partition oracle
(threads standing in for MPI ranks) agrees with the single-domain oracle where the algorithm is partition-independent,
and a world_size-2 gloo run exchanges matching halos."""
import os
import subprocess
import sys

import numpy as np
import pytest

from icsfoam_b200 import capi, cases
from oracle.pyoracle import Oracle, World
from tests.conftest import ROOT


def setup_world(case, n_parts, mode="x"):
    part, meshes = case.partition(n_parts, mode)
    w = World(n_parts)
    w.mesh_set(meshes)
    for r, o in enumerate(w.ranks):
        o.thermo_set(case.R, case.Cp, case.mu, case.Pr)
        o.schemes_set(case.schemes)
        names = [p["name"] for p in meshes[r].patches]
        for patch, fields in case.bcs.items():
            if patch in names:
                for field, (kind, params) in fields.items():
                    if isinstance(params, np.ndarray) and params.ndim == 2:   # non-uniform entries: the rows of this rank's faces
                        m = meshes[r]
                        params = np.ascontiguousarray(params[m.face_global[m.patch_faces(patch)] - case.mesh.patches[case.mesh.patch_index(patch)]["start"]])
                    o.bc_set(patch, {"p": 0, "U": 1, "T": 2}[field], kind, params)
        if case.mrf is not None:
            o.mrf_set(*case.mrf_fields(meshes[r]))
        if case.transport is not None:
            o.transport_set(*case.transport_fields(meshes[r]))
    w.state_set([case.p[m.cell_global] for m in meshes], [case.U[m.cell_global] for m in meshes], [case.T[m.cell_global] for m in meshes])
    return w, meshes


def test_partition_processor_patches_match():
    case = cases.onera_box(6)
    part, meshes = case.partition(4, (2, 2, 1))
    assert sum(m.n_cells for m in meshes) == case.mesh.n_cells
    for r, m in enumerate(meshes):
        assert (m.owner[: m.n_internal_faces] < m.neighbour).all()
        for p in m.patches:
            if p["kind"] != capi.PROCESSOR:
                continue
            other = meshes[p["nbr_rank"]]
            q = [x for x in other.patches if x["kind"] == capi.PROCESSOR and x["nbr_rank"] == r][0]
            assert q["size"] == p["size"]
            fa = np.arange(p["start"], p["start"] + p["size"])
            fb = np.arange(q["start"], q["start"] + q["size"])
            assert np.array_equal(m.face_global[fa], other.face_global[fb])       # same order on both sides
            assert np.allclose(m.Sf[fa], -other.Sf[fb], atol=0)                   # opposite orientation
            assert np.allclose(m.weights[fa] + other.weights[fb], 1.0, atol=1e-14)


@pytest.mark.parametrize("n_parts,mode,mu,mrf", [(2, "x", 0.0, False), (4, (2, 2, 1), 0.0, False), (4, (2, 2, 1), 0.5, False),
                                                 (4, (2, 2, 1), 0.0, True), (4, (2, 2, 1), 0.5, "transport"), (4, (2, 2, 1), 0.0, "profiles")])
def test_partitioned_oracle_matches_single_domain(n_parts, mode, mu, mrf):
    """Fluxes / residuals / SpMV do not depend on the decomposition (only LU-SGS and hence the GMRES history do —
    lusgs.C:149,181 keeps the sweeps rank-local).  mu > 0 adds the viscous residual, whose processor-patch faces need the
    neighbour's gradients of U and eCalc."""
    case = cases.onera_box(6, mu=mu)
    if mrf == "profiles":     # non-uniform freestream entries on the inlet patch, split over the ranks
        n = case.mesh.patches[case.mesh.patch_index("inlet")]["size"]
        y = case.mesh.Cf[case.mesh.patch_faces("inlet")][:, 1]
        case.bcs["inlet"] = dict(case.bcs["inlet"])
        case.bcs["inlet"]["U"] = ("freestream", np.ascontiguousarray(np.column_stack([285.6 * (1 + 0.02 * np.sin(y)), np.full(n, 15.268), np.zeros(n)])))
        case.bcs["inlet"]["T"] = ("inletOutlet", np.ascontiguousarray((288.15 + 3.0 * np.cos(y))[:, None]))
    elif mrf == "transport":  # muEff / alphaEff fields: processor halos carry the neighbour's cell values
        case.with_transport(lambda x: (0.5 * (1.2 + np.sin(2.0 * x[:, 0]) * np.cos(x[:, 2])), 0.8 * (1.1 + np.cos(1.5 * x[:, 1]))))
    elif mrf:  # rotating zone in half of the domain: MRFFaceVelocity on processor faces is each side's own (outward) value
        case.with_mrf(omega=(0.0, 40.0, 90.0), origin=(0.5, 0.0, 1.5), zone=lambda x: x[:, 0] > 0.2)
    single = case.apply(Oracle())
    phi, phiUp, phiEp = single.calc_flux()
    src = single.residual()
    w, meshes = setup_world(case, n_parts, mode)
    for r, (o, m) in enumerate(zip(w.ranks, meshes)):
        pass
    # run the flux on all ranks concurrently (halo exchange inside): one fused iteration drives it
    ctl = capi.solver_controls("Jacobi", n_directions=4, max_iter=30, tolerance=1e-14, rel_tol=1e-9)
    res_w = w.iterate(ctl, 1)
    res_s = single.iterate(ctl)
    # block-Jacobi preconditioning is decomposition independent -> same GMRES history up to reduction order
    assert res_w.n_iterations == res_s.n_iterations
    assert np.allclose(list(res_w.s_init) + list(res_w.v_init), list(res_s.s_init) + list(res_s.v_init), rtol=1e-9)
    st = single.state_get()
    for o, m in zip(w.ranks, meshes):
        sr = o.state_get()
        for k in ("rho", "rhoU", "rhoE"):
            assert np.abs(sr[k] - st[k][m.cell_global]).max() <= 1e-8 * np.abs(st[k]).max(), k


REF_VKI = "/root/reference/tutorials/VKI-LS89/constant/polyMesh"


@pytest.mark.skipif(not os.path.isdir(REF_VKI), reason="reference tutorial mesh not present on this machine")
def test_vki_ls89_decomposed_with_preserved_cyclic_pair():
    """C5 decomposed: the cyclic pair of the shipped VKI-LS89 mesh stays whole on each rank (decomposeParDict preservePatches,
    VKI-LS89/system/decomposeParDict:27-35); with block-Jacobi preconditioning the 2-rank world equals the single-domain run."""
    case = cases.vki_ls89(REF_VKI)
    ctl = capi.solver_controls("Jacobi", n_directions=8, max_iter=30, tolerance=1e-14, rel_tol=1e-8)
    single = case.apply(Oracle())
    w, meshes = setup_world(case, 2, "x")
    for m in meshes:
        up, lo = m.patches[m.patch_index("Upper_periodicity")], m.patches[m.patch_index("Lower_periodicity")]
        assert up["size"] == lo["size"] > 0 and up["kind"] == capi.CYCLIC
    for _ in range(2):
        rw, rs = w.iterate(ctl, 1), single.iterate(ctl)
        assert rw.n_iterations == rs.n_iterations
    st = single.state_get()
    for o, m in zip(w.ranks, meshes):
        sr = o.state_get()
        for k in ("rho", "rhoU", "rhoE"):
            assert np.abs(sr[k] - st[k][m.cell_global]).max() <= 1e-10 * np.abs(st[k]).max(), k


def test_lusgs_is_rank_local():
    """With LU-SGS the partitioned run is a different (block-Jacobi-of-LU-SGS) preconditioner: histories differ, results
    still converge to the same update within the linear tolerance."""
    case = cases.onera_box(6)
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=40, tolerance=1e-14, rel_tol=1e-8)
    single = case.apply(Oracle())
    single.iterate(ctl)
    w, meshes = setup_world(case, 2)
    w.iterate(ctl, 1)
    st = single.state_get()
    for o, m in zip(w.ranks, meshes):
        sr = o.state_get()
        assert np.abs(sr["rho"] - st["rho"][m.cell_global]).max() <= 1e-6 * np.abs(st["rho"]).max()


GLOO_SCRIPT = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["ICS_ROOT"])
from icsfoam_b200 import cases, capi
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = cases.onera_box(6)
part, meshes = case.partition(world, "x")
m = meshes[rank]
# every rank derives the same decomposition; exchange processor-patch face centres with the neighbour and compare
for p in m.patches:
    if p["kind"] != capi.PROCESSOR:
        continue
    f = np.arange(p["start"], p["start"] + p["size"])
    send = torch.from_numpy(np.ascontiguousarray(m.Cf[f]))
    recv = torch.empty_like(send)
    nb = p["nbr_rank"]
    reqs = [dist.isend(send, nb), dist.irecv(recv, nb)]
    [r.wait() for r in reqs]
    assert torch.allclose(send, recv, atol=0), "processor patch faces are not matched in order"
n = torch.tensor([m.n_cells], dtype=torch.int64)
dist.all_reduce(n)
assert int(n) == case.mesh.n_cells
print("rank", rank, "ok")
'''


def test_gloo_world_size_2_halo_consistency(tmp_path):
    script = tmp_path / "gloo_check.py"
    script.write_text(GLOO_SCRIPT)
    env = dict(os.environ, ICS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


GLOO_HB_SCRIPT = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["ICS_ROOT"])
from icsfoam_b200 import cases, capi
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# Harmonic Balance across ranks: every rank replicates ITS partition n_instants times; the processor patches of the
# replicated meshes must pair up instance by instance, in the same order on both sides (one NCCL group per exchange)
full = cases.hb_box(5, 3, flux="ROE", cyclic=False, seed=3)
mine = full.partition(world, "x")[rank]
m = mine.mesh                                  # hb.ReplicatedMesh of this rank's partition
procs = [p for p in m.patches if p["kind"] == capi.PROCESSOR]
assert len(procs) == 3 * (world - 1) and [p["name"].split("@")[1] for p in procs] == ["0", "1", "2"][: len(procs)]
for p in procs:
    f = np.arange(p["start"], p["start"] + p["size"])
    send = torch.from_numpy(np.ascontiguousarray(m.Cf[f]))
    recv = torch.empty_like(send)
    reqs = [dist.isend(send, p["nbr_rank"]), dist.irecv(recv, p["nbr_rank"])]
    [r.wait() for r in reqs]
    assert torch.allclose(send, recv, atol=0), "replicated processor patches are not matched in order"
    assert (m.owner[f] // mine.base.mesh.n_cells == int(p["name"].split("@")[1])).all()     # faces of instance K touch cells of instance K
n = torch.tensor([mine.base.mesh.n_cells], dtype=torch.int64)
dist.all_reduce(n)
assert int(n) == full.base.mesh.n_cells
print("rank", rank, "ok")
'''


def test_gloo_world_size_2_hb_replicated_partitions(tmp_path):
    script = tmp_path / "gloo_hb_check.py"
    script.write_text(GLOO_HB_SCRIPT)
    env = dict(os.environ, ICS_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT="29537")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29537", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_smooth_solver_is_partition_independent():
    """smoothSolverCoupled with the Jacobi smoother has no rank-local sweep: the 4-rank world follows the single domain sweep by sweep."""
    case = cases.onera_box(6)
    ctl = capi.solver_controls(solver="smoothSolverCoupled", n_sweeps=2, max_iter=8, tolerance=1e-14, rel_tol=1e-6)
    single = case.apply(Oracle())
    w, meshes = setup_world(case, 4, (2, 2, 1))
    for _ in range(2):
        rw, rs = w.iterate(ctl, 1), single.iterate(ctl)
        assert rw.n_iterations == rs.n_iterations == 8
        assert np.allclose(list(rw.s_final) + list(rw.v_final), list(rs.s_final) + list(rs.v_final), rtol=1e-9)
    st = single.state_get()
    for o, m in zip(w.ranks, meshes):
        sr = o.state_get()
        for k in ("rho", "rhoU", "rhoE"):
            assert np.abs(sr[k] - st[k][m.cell_global]).max() <= 1e-11 * np.abs(st[k]).max(), k
