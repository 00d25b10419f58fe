"""Multi-GPU parity driver (launched by torchrun, one rank per GPU): the P-GPU run must match the P-partition CPU
oracle with the same decomposition (LU-SGS is rank-local, so parity is defined per decomposition)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from icsfoam_b200 import capi, cases  # noqa: E402
from icsfoam_b200.context import Context  # noqa: E402


def main_hb(rank, world, local, ids):
    """Harmonic Balance on P GPUs: every rank holds all time instances of its partition; parity against the P-rank HB oracle world."""
    from oracle.pyoracle import HBWorld
    n_iter = int(os.environ.get("ICS_MULTI_ITERS", "4"))
    full = cases.hb_box(6, 3, flux="ROE", cyclic=False, seed=21)
    parts = full.partition(world, "x")
    mine = parts[rank]
    ctx = mine.apply(Context(device=local, nccl_id=ids[0], rank=rank, n_ranks=world))
    ctl = capi.solver_controls("LUSGS", n_directions=5, max_iter=10, tolerance=1e-10, rel_tol=1e-3)
    hist = []
    for _ in range(n_iter):
        r = ctx.iterate(ctl)
        hr = ctx.hb_residuals()
        hist.append(list(hr["s_init"]) + list(hr["v_init"]) + [r.n_iterations])
    st = ctx.state_get()
    payload = {"rho": st["rho"], "rhoU": st["rhoU"], "rhoE": st["rhoE"], "hist": hist}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    ok = True
    if rank == 0:
        W = HBWorld(parts)
        ohist = []
        for _ in range(n_iter):
            rr = W.iterate(ctl)[0]
            ohist.append(list(rr["s_init"]) + list(rr["v_init"]) + [rr["n_iterations"]])
        ohist, ghist = np.array(ohist), np.array(gathered[0]["hist"])
        print("oracle HB history", ohist[:, [0, 1, -1]].tolist())
        print("gpu    HB history", ghist[:, [0, 1, -1]].tolist())
        ok &= np.array_equal(ohist[:, -1], ghist[:, -1])
        ok &= np.allclose(ohist[:, :-1], ghist[:, :-1], rtol=1e-8, atol=1e-14)
        for r_, (h, g) in enumerate(zip(W.ranks, gathered)):
            so = h.state_get()
            for k in ("rho", "rhoU", "rhoE"):
                err = np.abs(g[k] - so[k]).max() / np.abs(so[k]).max()
                print(f"rank {r_} {k} rel err {err:.3e}")
                ok &= err <= 1e-8
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def decomposed_case(rank, world):
    """A warped hex box written as an OpenFOAM case and decomposed into processor<r> directories in decomposePar's layout
    (icsfoam_b200/meshtools/foamcase.py); every rank reads the directories back (variant "decomposed", not in the round-1 test list:
    added after the last multi-GPU call of the round)."""
    import tempfile
    from icsfoam_b200.meshtools import foamcase, read_polymesh
    box = [tempfile.mkdtemp(prefix="icsb200_decomposed_") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    case_dir = box[0]
    pts, faces, own, nei, patches = foamcase.box_polymesh(10, 8, 6, lo=(-1.0, 0.0, 0.0), hi=(2.0, 1.5, 1.0), warp=0.15)
    if rank == 0:
        foamcase.write_polymesh(os.path.join(case_dir, "constant", "polyMesh"), pts, faces, own, nei, patches)
    dist.barrier()
    whole = read_polymesh(os.path.join(case_dir, "constant", "polyMesh"))
    if rank == 0:
        order = np.argsort(whole.C[:, 0], kind="stable")
        part = np.empty(whole.n_cells, np.int32)
        part[order] = (np.arange(whole.n_cells) * world // whole.n_cells).astype(np.int32)
        foamcase.decompose(case_dir, pts, faces, own, nei, patches, part)
    dist.barrier()
    rng = np.random.default_rng(5)
    N = whole.n_cells
    p, T = 1e5 * (1 + 0.05 * rng.random(N)), 300.0 * (1 + 0.05 * rng.random(N))
    U = np.column_stack([120.0 + 10 * rng.random(N), 8 * rng.random(N), 5 * rng.random(N)])
    bcs = {q["name"]: {"p": ("zeroGradient", ()), "U": ("zeroGradient", ()), "T": ("zeroGradient", ())} for q in whole.patches}
    case = cases.Case("decomposed", whole, 287.0, 1005.0, capi.default_schemes(flux_scheme="ROE"),
                      capi.solver_controls("LUSGS", n_directions=5, max_iter=10, tolerance=1e-10, rel_tol=1e-3), bcs, p, U, T, mu=0.3)
    return case, case.decomposed(case_dir)[1]


def main_vki(rank, world, local, ids, prepared=None):
    """C5 across GPUs: the shipped VKI-LS89 mesh decomposed with its cyclic pair kept whole per rank (decomposeParDict
    preservePatches), laminar viscous ROE run; parity against the P-rank oracle world of the same decomposition."""
    from oracle.pyoracle import World
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n_iter = int(os.environ.get("ICS_MULTI_ITERS", "4"))
    if prepared is not None:
        case, meshes = prepared
    else:
        case = cases.vki_ls89(os.path.join(cases.tutorial_dir("VKI-LS89"), "constant", "polyMesh"))
        part, meshes = case.partition(world, "x")
    m = meshes[rank]
    ctx = case.apply(Context(device=local, nccl_id=ids[0], rank=rank, n_ranks=world), mesh=m, cells=m.cell_global)
    hist = []
    for _ in range(n_iter):
        r = ctx.iterate(case.controls)
        hist.append(list(r.s_init) + list(r.v_init)[:2] + [r.n_iterations])
    st = ctx.state_get()
    payload = {"rho": st["rho"], "rhoU": st["rhoU"], "rhoE": st["rhoE"], "hist": hist}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    ok = True
    if rank == 0:
        w = World(world)
        w.mesh_set(meshes)
        for o, mm in zip(w.ranks, meshes):
            o.thermo_set(case.R, case.Cp, case.mu, case.Pr)
            o.schemes_set(case.schemes)
            names = [p["name"] for p in mm.patches]
            for patch, fields in case.bcs.items():
                if patch in names:
                    for field, (kind, params) in fields.items():
                        o.bc_set(patch, {"p": 0, "U": 1, "T": 2}[field], kind, params)
        w.state_set([case.p[mm.cell_global] for mm in meshes], [case.U[mm.cell_global] for mm in meshes], [case.T[mm.cell_global] for mm in meshes])
        ohist = []
        for _ in range(n_iter):
            rr = w.iterate(case.controls, 1)
            ohist.append(list(rr.s_init) + list(rr.v_init)[:2] + [rr.n_iterations])
        ohist, ghist = np.array(ohist), np.array(gathered[0]["hist"])
        print("oracle VKI history", ohist[:, [0, 1, -1]].tolist())
        print("gpu    VKI history", ghist[:, [0, 1, -1]].tolist())
        ok &= np.array_equal(ohist[:, -1], ghist[:, -1])
        ok &= np.allclose(ohist[:, :-1], ghist[:, :-1], rtol=1e-8, atol=1e-14)
        for r_, (o, g) in enumerate(zip(w.ranks, gathered)):
            so = o.state_get()
            for k in ("rho", "rhoU", "rhoE"):
                err = np.abs(g[k] - so[k]).max() / np.abs(so[k]).max()
                print(f"rank {r_} {k} rel err {err:.3e}")
                ok &= err <= 1e-8
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def main():
    n = int(os.environ.get("ICS_MULTI_N", "10"))
    n_iter = int(os.environ.get("ICS_MULTI_ITERS", "4"))
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    parts = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
    mu = float(os.environ.get("ICS_MULTI_MU", "0"))   # > 0: laminar viscous residual (halo of eCalc and its gradient)
    ids = [Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    variant = os.environ.get("ICS_MULTI_VARIANT", "")   # "globaldt": non-local time stepping (gMax over ranks); "mrf": rotating zone; "hb"
    if variant == "hb":
        return main_hb(rank, world, local, ids)
    if variant == "vki":
        return main_vki(rank, world, local, ids)
    if variant == "decomposed":
        return main_vki(rank, world, local, ids, prepared=decomposed_case(rank, world))

    def make(r):
        c = cases.onera_box(n, parts=parts, rank=r, mu=mu)
        if variant == "globaldt":
            c.schemes.local_timestepping = 0
        if variant == "fullvisc":
            c.schemes.viscous_full_jacobian = 1
        if variant == "mrf":
            c.with_mrf(omega=(0.0, 40.0, 90.0), origin=(0.5, 0.0, 1.5), zone=lambda x: x[:, 0] > 0.2)
        return c

    case = make(rank)
    ctx = case.apply(Context(device=local, nccl_id=ids[0], rank=rank, n_ranks=world))
    hist = []
    flux0 = ctx.calc_flux()
    for _ in range(n_iter):
        r = ctx.iterate(case.controls)
        hist.append(list(r.s_init) + list(r.v_init) + [r.n_iterations])
    st = ctx.state_get()
    payload = {"rho": st["rho"], "rhoU": st["rhoU"], "rhoE": st["rhoE"], "phi": flux0[0], "hist": hist}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    ok = True
    if rank == 0:
        from oracle.pyoracle import World
        meshes = [make(r) for r in range(world)]
        w = World(world)
        w.mesh_set([c.mesh for c in meshes])
        for o, c in zip(w.ranks, meshes):
            o.thermo_set(c.R, c.Cp, c.mu, c.Pr)
            o.schemes_set(c.schemes)
            names = [p["name"] for p in c.mesh.patches]
            for patch, fields in c.bcs.items():
                if patch in names:
                    for field, (kind, params) in fields.items():
                        o.bc_set(patch, {"p": 0, "U": 1, "T": 2}[field], kind, params)
            if c.mrf is not None:
                o.mrf_set(*c.mrf_fields(c.mesh))
        w.state_set([c.p for c in meshes], [c.U for c in meshes], [c.T for c in meshes])
        ohist = []
        for _ in range(n_iter):
            r = w.iterate(case.controls, 1)
            ohist.append(list(r.s_init) + list(r.v_init) + [r.n_iterations])
        ohist, ghist = np.array(ohist), np.array(gathered[0]["hist"])
        print("oracle history", ohist[:, [0, 1, -1]].tolist())
        print("gpu    history", ghist[:, [0, 1, -1]].tolist())
        ok &= np.array_equal(ohist[:, -1], ghist[:, -1])
        ok &= np.allclose(ohist[:, :5], ghist[:, :5], rtol=1e-8, atol=1e-14)
        for r_, (o, g) in enumerate(zip(w.ranks, gathered)):
            so = o.state_get()
            for k in ("rho", "rhoU", "rhoE"):
                err = np.abs(g[k] - so[k]).max() / np.abs(so[k]).max()
                print(f"rank {r_} {k} rel err {err:.3e}")
                ok &= err <= 1e-8
        print("MULTI_GPU_PARITY", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
