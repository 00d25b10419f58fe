#!/usr/bin/env python
"""bench.py — Mcell-iterations/s of the ICSFoam implicit pseudo-time iteration on B200 (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload onera344|bump4m|forwardstep|vki|vki-hb]

A "step" is one outer pseudo-time iteration of dbnsFoam (outerLoop.H:51-99 + updateFields.H): gradients, flux,
residual, local pseudo time step, Jacobian assembly, GMRES(m)/LU-SGS solve, field update.
Default workload (the configuration BASELINE.json's metric is quoted on): C4, the synthetic OneraM6-scale 3-D transonic
mesh of SURVEY.md §8d (the shipped OneraM6 mesh is incomplete in the reference checkout), HLLC + vanLeer, steady, Co=100,
GMRES m=5 maxIter 10 relTol 0.1 with LU-SGS, at 344^3 = 40.7 M cells (~100 GB of HBM; --cells-per-dim to change).
Other named meshes (`--workload`): bump4m = C3 circularArcBump refined to 3 x 1280 x 1040 cells (HLLC, Minmod, Co 200,
m=5), forwardstep = C2, vki = C5 (i) and vki-hb = C5 (ii, Harmonic Balance with 3 time instances) on the reference's own shipped meshes.
Prints ONE JSON line.  `value` = device-resident throughput (inputs in HBM), `e2e` = the same iteration through
icsb200_iterate_host with pinned host buffers (p,U,T in and out every step).  `--impl reference` times the CPU
restatement of the same path (the reference itself needs OpenFOAM v2112, absent here) on all host cores on a bounded
sample of the same workload; both arms print the same `config`, the sample is described in `cpu_baseline.sample`.
With N > 1 ranks a parity pre-flight runs first: the same decomposition at a small size on the P GPUs and in the P-rank
CPU oracle world (the checker, never the thing measured); a miss ends the run with a non-zero exit code.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

_ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, _ROOT)

METRIC = "Mcell-iterations/s (flux+Jacobian+GMRES)"
UNIT = "Mcell-it/s"
GPU_PARTS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
# fp64 operation slots per face of the flux kernel (DESIGN.md section 4 "Face kernels": 42 IEEE divisions / square roots at
# ~14 multiply slots each + ~400 other operations) and the measured fp64 issue rate (profiles/r01_fp64_latency.txt)
FLUX_SLOTS_PER_FACE = 1000.0
FP64_SLOTS_PER_S = 61.0 * 148 * 1.965e9


def peaks():
    p = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def factor3(p):
    """p = px * py * pz with the factors as equal as possible (px >= py >= pz)."""
    best = (p, 1, 1)
    for a in range(1, p + 1):
        if p % a:
            continue
        for b in range(1, p // a + 1):
            if (p // a) % b:
                continue
            c = p // a // b
            t = tuple(sorted((a, b, c), reverse=True))
            if max(t) - min(t) < max(best) - min(best):
                best = t
    return best


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
class Workload:
    """One named mesh + solver set-up.  `gpu(world, rank)` -> (case, mesh of this rank or None, global cell ids or None);
    `cpu(threads)` -> (list of (case, mesh, global cell ids or None) per partition, cells, description of the bounded sample)."""

    def __init__(self, args):
        self.args = args
        from icsfoam_b200 import cases
        self.cases = cases
        w = args.workload
        if w == "onera344":
            n = args.n
            self.n_total = n ** 3
            self.label = (f"onera-box {n}^3 = {n**3} cells (C4 synthetic OneraM6-scale 3-D transonic), HLLC vanLeer steady Co=100, "
                          f"GMRES m=5 maxIter 10 relTol 0.1, LU-SGS")
        elif w == "bump4m":
            self.n_total = 3 * args.bump_nx * args.bump_ny
            self.label = (f"circularArcBump transonic 3x{args.bump_nx}x{args.bump_ny} = {self.n_total} cells (C3, blockMesh-refined 2-D), "
                          f"HLLC Minmod steady Co=200, GMRES m=5 maxIter 10 relTol 1e-2, LU-SGS")
        elif w in ("forwardstep", "vki", "vki-hb"):
            name = {"forwardstep": "forwardStep", "vki": "VKI-LS89", "vki-hb": "VKI-LS89"}[w]
            d = cases.tutorial_dir(name)
            if d is None:
                raise SystemExit(f"workload {w}: tutorial {name} not found ($ICSFOAM_REF, /root/reference or cases_local/)")
            self.poly = os.path.join(d, "constant", "polyMesh")
            self._whole = (cases.forward_step(self.poly) if w == "forwardstep" else cases.vki_ls89(self.poly) if w == "vki"
                           else cases.vki_hb(self.poly, 3))
            self.n_total = self._whole.mesh.n_cells   # Harmonic Balance: the coupled cells of all time instances
            self.label = {"forwardstep": f"forwardStep Mach 3 shipped polyMesh = {self.n_total} cells (C2), HLLC Minmod, dual time (backward), GMRES m=8 maxIter 20 relTol 1e-4, LU-SGS",
                          "vki": f"VKI-LS89 shipped polyMesh = {self.n_total} cells (C5 i, cyclic pair, laminar viscous), ROE vanLeer steady Co=10, GMRES m=8 maxIter 10 relTol 1e-3, LU-SGS",
                          "vki-hb": f"VKI-LS89 Harmonic Balance, 3 time instances x 28059 cells = {self.n_total} coupled cells (C5 ii, dbnsFullyImplicitHBFoam; inlet total pressure "
                                    f"oscillating at 2 kHz), ROE vanLeer Co=10, GMRES m=8 maxIter 10 relTol 1e-3, LU-SGS"}[w]
        else:
            raise SystemExit(f"unknown workload {w}")
        small = self.n_total * 2600 < 4 * 126e6
        self.flush_l2 = small
        self.config = {"workload": self.label, "cells": int(self.n_total),
                       "l2": ("L2 flushed (256 MB written) between timed steps, each step timed on its own" if small
                              else "working set per step >> 126 MB L2 (inputs larger than L2)")}

    def _whole_case(self, sample=False):
        a, c = self.args, self.cases
        if a.workload == "bump4m":
            return c.bump(a.bump_nx_cpu, a.bump_ny_cpu) if sample else c.bump(a.bump_nx, a.bump_ny)
        return self._whole

    def gpu(self, world, rank):
        a, c = self.args, self.cases
        if a.workload == "onera344":
            if world == 1:
                return c.onera_box(a.n), None, None
            # strong scaling: the SAME mesh split into `world` blocks (decomposePar-style processor patches); every rank
            # generates only its own block (+ one ghost layer), never the global mesh
            parts = GPU_PARTS.get(world, (world, 1, 1))
            return c.onera_box(a.n, parts=parts, rank=rank), None, None
        case = self._whole_case()
        if world == 1:
            return case, None, None
        if a.workload == "vki-hb":   # every rank holds all time instances of its own partition (HB instants are not sharded)
            return case.partition(world, "x")[rank], None, None
        part, meshes = case.partition(world, "x", only=rank)
        return case, meshes[rank], meshes[rank].cell_global

    def cpu(self, threads):
        a, c = self.args, self.cases
        if a.workload == "onera344":
            px = factor3(threads)
            parts = [(c.onera_box(a.n_cpu, parts=px, rank=r), None, None) for r in range(threads)] if threads > 1 else [(c.onera_box(a.n_cpu), None, None)]
            return parts, a.n_cpu ** 3, f"onera-box {a.n_cpu}^3 ({a.n_cpu**3} cells) in {px[0]}x{px[1]}x{px[2]} blocks"
        case = self._whole_case(sample=True)
        if a.workload == "vki-hb":
            threads = max(1, min(threads, 8))
            return case.partition(threads, "x"), case.mesh.n_cells, f"{case.name} {case.mesh.n_cells} coupled cells in {threads} x-slabs (all instances per slab)"
        threads = max(1, min(threads, case.mesh.n_cells // 2000))
        if threads == 1:
            return [(case, None, None)], case.mesh.n_cells, f"{case.name} {case.mesh.n_cells} cells, 1 partition"
        part, meshes = case.partition(threads, "x")
        return [(case, m, m.cell_global) for m in meshes], case.mesh.n_cells, f"{case.name} {case.mesh.n_cells} cells in {threads} x-slabs"


def apply_to_world(parts):
    """P oracle contexts on P threads (mailbox halo exchange, rank-ordered reductions): the stand-in for the reference's MPI run."""
    from oracle.pyoracle import World
    w = World(len(parts))
    meshes = [m if m is not None else c.mesh for c, m, g in parts]
    w.mesh_set(meshes)
    fid = {"p": 0, "U": 1, "T": 2}
    for o, (c, m, g), mesh in zip(w.ranks, parts, meshes):
        o.thermo_set(c.R, c.Cp, c.mu, c.Pr)
        o.schemes_set(c.schemes)
        names = [p["name"] for p in mesh.patches]
        for patch, fields in c.bcs.items():
            if patch in names:
                for field, (kind, params) in fields.items():
                    o.bc_set(patch, fid[field], kind, params)
    sel = lambda c, g, x: x if g is None else x[g]
    w.state_set([sel(c, g, c.p) for c, m, g in parts], [sel(c, g, c.U) for c, m, g in parts], [sel(c, g, c.T) for c, m, g in parts])
    return w


class CpuRun:
    """The CPU restatement (oracle 'port') on the host cores: P partitions on P threads, mirroring P MPI ranks
    (halo exchange between threads, rank-ordered reductions, LU-SGS local to each partition as lusgs.C:149,181)."""

    def __init__(self, wl, threads):
        from oracle.pyoracle import Oracle
        parts, self.n_cells, self.sample = wl.cpu(threads)
        self.threads = len(parts)
        self.hb = None
        if wl.args.workload == "vki-hb":      # the reference-structured HB oracle: nO contexts per rank + the global (2 nO, nO) system
            from oracle.pyoracle import HBWorld
            self.controls = parts[0].controls
            self.hb = HBWorld(parts)
            return
        self.controls = parts[0][0].controls
        if self.threads > 1:
            self.world = apply_to_world(parts)
        else:
            self.single = parts[0][0].apply(Oracle())

    def iterate(self, iters):
        t0 = time.perf_counter()
        if self.hb is not None:
            n_it = self.hb.iterate(self.controls, iters)[0]["n_iterations"]
            dt = time.perf_counter() - t0
            return self.n_cells * iters / dt / 1e6, dt, n_it
        if self.threads > 1:
            res = self.world.iterate(self.controls, iters)
        else:
            for _ in range(iters):
                res = self.single.iterate(self.controls)
        dt = time.perf_counter() - t0
        return self.n_cells * iters / dt / 1e6, dt, res.n_iterations


def host_threads():
    """All host cores (the reference would run one MPI rank per core)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's own CPU path cannot be built here (OpenFOAM v2112 absent), so this arm times the
    CPU restatement with all host threads it can use; each step is one outer iteration of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = Workload(args)
    run = CpuRun(wl, host_threads())
    for _ in range(args.warmup):
        run.iterate(1)
    t_all, restarts = 0.0, []
    for _ in range(args.steps):
        v, dt, r = run.iterate(1)
        t_all += dt; restarts.append(r)
    value = run.n_cells * args.steps / t_all / 1e6
    sample = (f"{run.sample}, 1 outer iteration per step, {run.threads} partitions on {run.threads} threads; CPU restatement of the "
              f"ICSFoam path (oracle/, -O3), not the OpenFOAM binary")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "config": wl.config,
            "run": {"restarts_per_step": float(np.mean(restarts)), "sample_cells": int(run.n_cells),
                    "note": "bounded sample of config.workload: same mesh family, schemes and solver controls at the size stated in cpu_baseline.sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": run.threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def preflight_parity(world, rank, local_rank, dist, torch, n=24, iters=3):
    """The same block decomposition at n^3 cells on the P GPUs and in the P-rank oracle world (blockFvMatrixUpdateMatrixInterfaces.C:33-185
    halo semantics, lusgs.C:149,181 rank-local sweeps): residual history, restart counts and state, tests/common.py bars."""
    from icsfoam_b200 import cases
    from icsfoam_b200.context import Context
    ids = [Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    parts = GPU_PARTS.get(world, (world, 1, 1))
    case = cases.onera_box(n, parts=parts, rank=rank)
    ctx = case.apply(Context(device=local_rank, nccl_id=ids[0], rank=rank, n_ranks=world))
    hist = []
    for _ in range(iters):
        r = ctx.iterate(case.controls)
        hist.append(list(r.s_init) + list(r.v_init) + [r.n_iterations])
    st = ctx.state_get()
    payload = {"rho": st["rho"], "rhoU": st["rhoU"], "rhoE": st["rhoE"], "hist": hist}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    out = None
    if rank == 0:
        w = apply_to_world([(cases.onera_box(n, parts=parts, rank=r), None, None) for r in range(world)])
        ohist = []
        for _ in range(iters):
            r = w.iterate(case.controls, 1)
            ohist.append(list(r.s_init) + list(r.v_init) + [r.n_iterations])
        ohist, ghist = np.array(ohist), np.array(gathered[0]["hist"])
        restarts_equal = bool(np.array_equal(ohist[:, -1], ghist[:, -1]))
        hist_rel = float(np.max(np.abs(ohist[:, :5] - ghist[:, :5]) / np.maximum(np.abs(ohist[:, :5]), 1e-300)))
        max_rel = 0.0
        for o, g in zip(w.ranks, gathered):
            so = o.state_get()
            for k in ("rho", "rhoU", "rhoE"):
                max_rel = max(max_rel, float(np.abs(g[k] - so[k]).max() / np.abs(so[k]).max()))
        ok = restarts_equal and hist_rel <= 1e-8 and max_rel <= 1e-8
        out = {"ranks": world, "cells": n ** 3, "outer_iterations": iters, "max_rel": max_rel, "residual_history_max_rel": hist_rel,
               "restarts_equal": restarts_equal, "ok": bool(ok), "checker": "P-rank CPU oracle world, same decomposition"}
    flag = torch.tensor([1 if (out is None or out["ok"]) else 0], device="cuda")
    dist.broadcast(flag, src=0)
    ctx.close()
    return out, int(flag.item()) == 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default=os.environ.get("ICSB200_BENCH_WORKLOAD", "onera344"), choices=["onera344", "bump4m", "forwardstep", "vki", "vki-hb"])
    ap.add_argument("--cells-per-dim", "--n", dest="n", type=int, default=int(os.environ.get("ICSB200_BENCH_N", "344")),
                    help="cells per direction of the 3-D mesh (use the long form under torchrun)")
    ap.add_argument("--cpu-cells-per-dim", "--n-cpu", dest="n_cpu", type=int, default=128, help="cells per direction of the bounded CPU-baseline sample")
    ap.add_argument("--bump-nx", type=int, default=1280, help="cells per block and direction x of the bump mesh (3 blocks)")
    ap.add_argument("--bump-ny", type=int, default=1040)
    ap.add_argument("--bump-nx-cpu", type=int, default=640)
    ap.add_argument("--bump-ny-cpu", type=int, default=520)
    ap.add_argument("--second-co", type=float, default=10.0, help="pseudo-Courant number of the second, converging solver regime (0 = skip)")
    ap.add_argument("--skip-cpu", "--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--skip-e2e", "--no-e2e", dest="no_e2e", action="store_true")
    ap.add_argument("--skip-extra", dest="no_extra", action="store_true", help="skip the second regime and the run at the CPU sample's size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        args.gpus = world
    import torch
    import torch.distributed as dist
    from icsfoam_b200.context import Context
    multi = world > 1
    if multi:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    parity = None
    if multi:
        parity, ok = preflight_parity(world, rank, local_rank, dist, torch)
        if not ok:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "error": "multi-GPU parity pre-flight failed",
                                  "parity": parity}), flush=True)
            dist.barrier()
            dist.destroy_process_group()
            sys.exit(3)

    wl = Workload(args)
    nccl_id = None
    if multi:
        ids = [Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
    case, mesh, cells = wl.gpu(world, rank)
    N_total = wl.n_total
    ctx = Context(device=local_rank, nccl_id=nccl_id, rank=rank, n_ranks=world) if multi else Context(device=local_rank)
    case.apply(ctx, mesh=mesh, cells=cells) if mesh is not None else case.apply(ctx)
    ctl = case.controls
    sched = ctx.schedule_info()

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device="cuda") if wl.flush_l2 else None

    def timed_steps(k):
        """k outer iterations timed with CUDA events on the library's stream; small workloads: every step on its own, L2 flushed between."""
        rs = []
        if flush_buf is None:
            ctx.timer_begin()
            for _ in range(k):
                rs.append(ctx.iterate(ctl).n_iterations)
            return ctx.timer_end(), rs
        ms = 0.0
        for _ in range(k):
            flush_buf.fill_(1.0)
            torch.cuda.synchronize()
            ctx.timer_begin()
            rs.append(ctx.iterate(ctl).n_iterations)
            ms += ctx.timer_end()
        return ms, rs

    # ---- device-resident timing
    for _ in range(args.warmup):
        ctx.iterate(ctl)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    barrier()
    ms, restarts = timed_steps(args.steps)
    barrier()
    launches = ctx.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if multi:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = N_total * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel-class device times (CUDA events on the launching stream) for the roofline of the dominant kernel
    ctx.timers_reset(True)
    prof_steps = min(3, args.steps)
    prof_restarts = [ctx.iterate(ctl).n_iterations for _ in range(prof_steps)]
    timers = ctx.timers_get()
    ctx.timers_reset(False)
    Nl, Fl = ctx.mesh.n_cells, ctx.mesh.n_internal_faces
    alg = {"spmv": 280 * Nl + 408 * Fl, "lusgs": 176 * Nl + 416 * Fl, "gradient": 264 * Nl + 40 * Fl,
           "flux_residual": 320 * Nl + 72 * Fl, "jacobian": (520 - 320 + 264 - 64) * Nl + (472 - 72 + 72) * Fl}
    total_ms = sum(v[0] for v in timers.values())
    dom = max((k for k in timers if k in alg), key=lambda k: timers[k][0])
    peak, peak_kind = peaks()
    avg_ms = timers[dom][0] / max(timers[dom][1], 1)
    achieved = alg[dom] / (avg_ms * 1e-3) / 1e9
    breakdown = {}
    for k, v in timers.items():
        if not v[1]:
            continue
        e = {"ms_per_step": round(v[0] / prof_steps, 4), "launches_per_step": v[1] / prof_steps, "share": round(v[0] / total_ms, 4)}
        if k in alg:
            gbs = alg[k] / (v[0] / v[1] * 1e-3) / 1e9
            e.update({"GBps": round(gbs, 1), "bound": "hbm", "frac": round(gbs / peak, 4)})
        if k == "flux_residual":
            # the flux evaluation is bound by the fp64 pipe, not by HBM: ~1000 operation slots per face
            t_fp64 = FLUX_SLOTS_PER_FACE * Fl / FP64_SLOTS_PER_S * 1e3
            e.update({"bound": "fp64", "frac": round(t_fp64 / (v[0] / v[1]), 4), "hbm_frac": e["frac"], "fp64_bound_ms": round(t_fp64, 3)})
        breakdown[k] = e
    # DRAM traffic of the dominant kernel: recorded from one `ncu --set full` capture of this very workload
    # (profiles/traffic.json; ncu cannot run inside the timed bench), null for any other workload / size / partition
    traffic = None
    try:
        rec = json.load(open(os.path.join(_ROOT, "profiles", "traffic.json")))
        key = f"onera-box-{args.n}-{world}gpu" if args.workload == "onera344" else f"{args.workload}-{world}gpu"
        traffic = rec.get(key, {}).get(dom)
    except Exception:
        traffic = None
    step_alg = (alg["gradient"] + alg["flux_residual"] + alg["jacobian"] + alg["spmv"] + 160 * Nl
                + float(np.mean(prof_restarts)) * ((ctl.n_directions + 1) * (alg["spmv"] + alg["lusgs"])
                                                   + Nl * (100 * ctl.n_directions ** 2 + 260 * ctl.n_directions + 200)))
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_kind,
                "whole_step_frac": round(step_alg / (ms / args.steps * 1e-3) / 1e9 / peak, 4),
                "algorithmic_bytes_per_launch": alg[dom], "avg_launch_ms": round(avg_ms, 4), "kernel_classes": breakdown}

    # ---- end to end through host buffers (pinned), p,U,T in and out every step
    e2e = None
    if not args.no_e2e:
        # every rank feeds its own partition from its own pinned buffers (iterate_host is collective like iterate: halo
        # exchanges and reductions inside); wall clock between barriers, max over ranks
        st = ctx.state_get()
        Nc = ctx.mesh.n_cells
        hp = torch.empty(Nc, dtype=torch.float64).pin_memory()
        hU = torch.empty((Nc, 3), dtype=torch.float64).pin_memory()
        hT = torch.empty(Nc, dtype=torch.float64).pin_memory()
        hp.numpy()[:] = st["p"]; hU.numpy()[:] = st["U"]; hT.numpy()[:] = st["T"]
        e_steps = max(3, args.steps // 2)
        e_err = None
        try:
            for _ in range(2):
                ctx.iterate_host(ctl, hp.numpy(), hU.numpy(), hT.numpy())
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                ctx.iterate_host(ctl, hp.numpy(), hU.numpy(), hT.numpy())
            barrier()
            dt = time.perf_counter() - t0
        except Exception as ex:      # reported, never silently replaced by the device-resident number
            e_err, dt = str(ex), float("inf")
        nbytes = 40 * Nc
        if multi:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
            b = torch.tensor([nbytes], device="cuda", dtype=torch.int64)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            nbytes = int(b.item())
        e2e = {"value": N_total * e_steps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
               "steps": e_steps}
        if e_err is not None or not np.isfinite(dt):
            e2e = {"value": None, "unit": UNIT, "error": e_err or "failed on another rank"}

    # ---- a second solver regime of the same workload: lower pseudo-Courant number, GMRES reaches relTol within 1-2 restarts,
    # so assembly (gradients, flux, Jacobian) carries weight (BASELINE.md section 3 quotes r = 1)
    regimes = None
    if not args.no_extra and args.second_co > 0:
        sch = case.schemes
        co0, cm0 = sch.pseudo_co_num, sch.pseudo_co_num_max
        sch.pseudo_co_num = sch.pseudo_co_num_max = args.second_co
        ctx.schemes_set(sch)
        for _ in range(max(3, args.warmup)):
            ctx.iterate(ctl)
        barrier()
        k2 = max(5, args.steps // 2)
        ms2, r2 = timed_steps(k2)
        barrier()
        if multi:
            t = torch.tensor([ms2], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms2 = float(t.item())
        regimes = [{"pseudoCoNum": args.second_co, "steps": k2, "restarts_per_step": float(np.mean(r2)), "ms_per_step": ms2 / k2,
                    "value": N_total * k2 / (ms2 * 1e-3) / 1e6, "unit": UNIT}]
        sch.pseudo_co_num, sch.pseudo_co_num_max = co0, cm0

    if rank != 0:
        if multi:
            dist.barrier()
            ctx.close()
            dist.destroy_process_group()
        return
    cpu, same_size = None, None
    if not args.no_cpu and not multi:
        run = CpuRun(wl, host_threads())
        run.iterate(1)
        v, dt, r = run.iterate(3)
        cpu = {"value": v, "unit": UNIT, "cores": run.threads, "kind": "port",
               "sample": f"{run.sample}, 3 outer iterations ({r} restarts in the last), {run.threads} partitions on "
                         f"{run.threads} threads, {dt:.1f} s; CPU restatement of the ICSFoam path, not the OpenFOAM binary"}
        if not args.no_extra and args.workload == "onera344" and args.n != args.n_cpu:
            # the GPU at the CPU sample's own size, so the two can be compared like for like (the headline stays the named size)
            from icsfoam_b200 import cases
            c2 = cases.onera_box(args.n_cpu)
            g2 = c2.apply(Context(device=local_rank))
            for _ in range(3):
                g2.iterate(c2.controls)
            g2.timer_begin()
            for _ in range(5):
                g2.iterate(c2.controls)
            ms3 = g2.timer_end()
            same_size = {"cells": args.n_cpu ** 3, "value": args.n_cpu ** 3 * 5 / (ms3 * 1e-3) / 1e6, "unit": UNIT, "steps": 5,
                         "note": "this library on one GPU at the size of cpu_baseline.sample (device-resident)"}
            g2.close()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": wl.config,
            "run": {"restarts_per_step": float(np.mean(restarts)), "lusgs_levels": sched["n_levels_fwd"],
                    "lusgs_schedule": ("block tiles: %d tiles, %d tile levels" % (sched["n_tiles"], sched["n_tile_levels"])) if sched.get("blk") else "level pipeline",
                    "partition": "1" if not multi else f"{world} blocks (processor patches)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    if regimes:
        line["regimes"] = regimes
    if same_size:
        line["gpu_at_cpu_sample_size"] = same_size
    if parity:
        line["parity"] = parity
    print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
