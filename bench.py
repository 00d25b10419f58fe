#!/usr/bin/env python
"""bench.py — Mcell-iterations/s of the ICSFoam implicit pseudo-time iteration on B200 (BASELINE.json metric).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--cells-per-dim N]

A "step" is one outer pseudo-time iteration of dbnsFoam (outerLoop.H:51-99 + updateFields.H): gradients, flux,
residual, local pseudo time step, Jacobian assembly, GMRES(m)/LU-SGS solve, field update.
Workload at N=1: the synthetic OneraM6-scale 3-D transonic mesh of SURVEY.md §8d (config C4; the shipped OneraM6
mesh is incomplete in the reference checkout), HLLC + vanLeer, steady, Co=100, GMRES m=5 maxIter 10 relTol 0.1 with
LU-SGS, at 344^3 = 40.7 M cells (the OneraM6-scale size of BASELINE.json; ~100 GB of HBM; --cells-per-dim to change).
Prints ONE JSON line.  `value` = device-resident throughput (inputs in HBM), `e2e` = the same iteration through
icsb200_iterate_host with pinned host buffers (p,U,T in and out every step).
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


def peaks():
    p = os.path.join(_ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


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


def make_case(n):
    from icsfoam_b200 import cases
    return cases.onera_box(n)


class CpuRun:
    """The CPU restatement (oracle 'port') on the host cores: P partitions on P threads, mirroring P MPI ranks
    (halo exchange between threads, rank-ordered reductions, LU-SGS local to each partition as lusgs.C:149,181)."""

    def __init__(self, n_cpu, threads):
        from icsfoam_b200 import cases
        from oracle.pyoracle import Oracle, World
        self.threads = threads
        if threads > 1:
            px = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2), 16: (4, 2, 2), 32: (4, 4, 2)}.get(threads, (threads, 1, 1))
            parts = [cases.onera_box(n_cpu, parts=px, rank=r) for r in range(threads)]
            self.case = parts[0]
            self.world = World(threads)
            self.world.mesh_set([c.mesh for c in parts])
            for o, c in zip(self.world.ranks, parts):
                o.thermo_set(c.R, c.Cp, c.mu, c.Pr)
                o.schemes_set(c.schemes)
                names = [p["name"] for p in c.mesh.patches]
                for patch, fields in c.bcs.items():
                    if patch in names:
                        for field, (kind, params) in fields.items():
                            o.bc_set(patch, {"p": 0, "U": 1, "T": 2}[field], kind, params)
            self.world.state_set([c.p for c in parts], [c.U for c in parts], [c.T for c in parts])
            self.n_cells = n_cpu ** 3
        else:
            self.case = cases.onera_box(n_cpu)
            self.single = self.case.apply(Oracle())
            self.n_cells = self.case.mesh.n_cells

    def iterate(self, iters):
        t0 = time.perf_counter()
        if self.threads > 1:
            res = self.world.iterate(self.case.controls, iters)
        else:
            for _ in range(iters):
                res = self.single.iterate(self.case.controls)
        dt = time.perf_counter() - t0
        return self.n_cells * iters / dt / 1e6, dt, res.n_iterations


def host_threads():
    cores = os.cpu_count() or 1
    t = 1
    while t * 2 <= min(cores, 32):
        t *= 2
    return t


def run_reference(args):
    """--impl reference: the reference's own CPU path cannot be built here (OpenFOAM v2112 absent), so this arm times the
    CPU restatement with all host threads it can use; each step is one outer iteration of a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    run = CpuRun(args.n_cpu, threads)
    for _ in range(args.warmup):
        run.iterate(1)
    vals, t_all, restarts = [], 0.0, []
    for _ in range(args.steps):
        v, dt, r = run.iterate(1)
        vals.append(v); t_all += dt; restarts.append(r)
    value = run.n_cells * args.steps / t_all / 1e6
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * t_all / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"onera-box {args.n}^3 (C4 synthetic OneraM6-scale 3-D transonic), HLLC vanLeer steady Co=100, GMRES m=5 "
                                   f"maxIter 10 relTol 0.1, LU-SGS",
                       "restarts_per_step": float(np.mean(restarts)),
                       "note": "CPU restatement of the ICSFoam path (OpenFOAM v2112 cannot be built here); bounded sample of the workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"onera-box {args.n_cpu}^3 ({args.n_cpu**3} cells), 1 outer iteration per step, {threads} partitions "
                                       f"on {threads} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--cells-per-dim", "--n", dest="n", type=int, default=int(os.environ.get("ICSB200_BENCH_N", "344")),
                    help="cells per direction of the 3-D mesh (use the long form under torchrun)")
    ap.add_argument("--cpu-cells-per-dim", "--n-cpu", dest="n_cpu", type=int, default=128, help="cells per direction of the bounded CPU-baseline sample")
    ap.add_argument("--skip-cpu", "--no-cpu", dest="no_cpu", action="store_true")
    ap.add_argument("--skip-e2e", "--no-e2e", dest="no_e2e", action="store_true")
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

    nccl_id = None
    if multi:
        from icsfoam_b200 import cases
        ids = [Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        nccl_id = ids[0]
        # strong scaling: the SAME mesh split into `world` blocks (decomposePar-style processor patches); every rank
        # generates only its own block (+ one ghost layer), never the global mesh
        parts = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}.get(world, (world, 1, 1))
        case = cases.onera_box(args.n, parts=parts, rank=rank)
        N_total = args.n ** 3
        ctx = Context(device=local_rank, nccl_id=nccl_id, rank=rank, n_ranks=world)
        case.apply(ctx)
    else:
        case = make_case(args.n)
        N_total = case.mesh.n_cells
        ctx = Context(device=local_rank)
        case.apply(ctx)
    ctl = case.controls
    sched = ctx.schedule_info()

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    restarts = []
    for _ in range(args.warmup):
        ctx.iterate(ctl)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    barrier()
    ctx.timer_begin()
    for _ in range(args.steps):
        restarts.append(ctx.iterate(ctl).n_iterations)
    ms = ctx.timer_end()
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
    breakdown = {k: {"ms_per_step": round(v[0] / prof_steps, 4), "launches_per_step": v[1] / prof_steps,
                     "share": round(v[0] / total_ms, 4),
                     **({"GBps": round(alg[k] / (v[0] / max(v[1], 1) * 1e-3) / 1e9, 1)} if k in alg and v[1] else {})}
                 for k, v in timers.items() if v[1]}
    # DRAM traffic of the dominant kernel: recorded from one `ncu --set full` capture of this very workload
    # (profiles/traffic.json; ncu cannot run inside the timed bench), null for any other size / partition
    traffic = None
    try:
        rec = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")))
        traffic = rec.get(f"onera-box-{args.n}-{world}gpu", {}).get(dom)
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

    if rank != 0:
        if multi:
            dist.barrier()
            ctx.close()
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu and not multi:
        threads = host_threads()
        run = CpuRun(args.n_cpu, threads)
        run.iterate(1)
        v, dt, r = run.iterate(3)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"onera-box {args.n_cpu}^3 ({args.n_cpu**3} cells), 3 outer iterations ({r} restarts in the last), {threads} partitions on "
                         f"{threads} threads, {dt:.1f} s; CPU restatement of the ICSFoam path, not the OpenFOAM binary"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"onera-box {args.n}^3 = {N_total} cells (C4 synthetic OneraM6-scale 3-D transonic), HLLC vanLeer steady "
                                   f"Co=100, GMRES m=5 maxIter 10 relTol 0.1, LU-SGS (level-scheduled, reference cell order)",
                       "restarts_per_step": float(np.mean(restarts)), "lusgs_levels": sched["n_levels_fwd"],
                       "l2": "working set per step >> 126 MB L2 (inputs larger than L2)", "partition": "1" if not multi else f"{world} blocks"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
