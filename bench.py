#!/usr/bin/env python
"""bench.py -- lattice site-updates/s of the Langevin hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

One "step" = one reference-style frame (tauhost.c:479-560): `loops` tau-steps of the fused
noise+stencil+update kernel over the whole lattice plus the per-step observable reductions.
Default at N=1: configs[1] of BASELINE.json -- 2-D 1024^2, fp32, dtau=0.01, cold start,
seed 1242608872; K=10 steps x loops=1000 = the config's 10^4 tau-steps.

Prints ONE JSON line (rank 0).  Timing: CUDA events on the stream the kernels are launched on
(the library's own stream, wrapped as a torch ExternalStream), barrier + synchronize on both
sides, max over ranks.  `value`: state resident in HBM.  `e2e`: the same frames through
sq_frame_host with pinned HOST buffers (H2D of the field before, D2H of field + observables
after, inside the timed region).  `roofline`: algorithmic bytes (8 B per fp32 site-update,
SURVEY.md 8(d)) over the update kernel's average launch duration (sq_kernel_timing: CUDA events
around every launch) against MEASURED_PEAKS.json's hbm_gbs.  `cpu_baseline`: the oracle's
OpenMP port timed on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: dims, real, potential, m2, lam, dtau, loops (tau-steps per bench step), math
    "c2": dict(dims=(1024, 1024), real="f32", pot=0, m2=0.0, lam=0.0, dtau=0.01, loops=1000,
               desc="configs[1]: 2-D 1024^2 fp32, 10^4 tau-steps = 10 steps x 1000"),
    "c2phi4": dict(dims=(1024, 1024), real="f32", pot=4, m2=0.25, lam=0.5, dtau=0.01, loops=1000,
                   desc="configs[1] with the phi^4 force (potID 4: m2=0.25, lambda=0.5)"),
    "c3": dict(dims=(64, 64, 64, 64), real="f32", pot=0, m2=0.0, lam=0.0, dtau=0.01, loops=100,
               desc="configs[2]: 4-D 64^4 fp32, 100 tau-steps per step"),
    "slab": dict(dims=(256, 256, 256, 32), real="f32", pot=0, m2=0.0, lam=0.0, dtau=0.01, loops=10,
                 desc="configs[3] per-GPU slab: 256^3 x 32 fp32 (x8 GPUs = 256^4)"),
    # configs[4]: independent chains over a coupling grid; 512 chains per GPU (x8 GPUs = 4096)
    "c5": dict(dims=(32, 32, 32, 32), real="f32", pot=4, m2=0.25, lam=0.5, dtau=0.01, loops=100, nchains=512,
               desc="configs[4] per-GPU share: 512 independent 32^4 chains, lambda in linspace(0,1,64) x 8 seeds, 100 tau-steps per step"),
    # one lattice, slab-decomposed along the time axis over the N ranks (strong scaling)
    "c4": dict(dims=(256, 256, 256, 256), real="f32", pot=0, m2=0.0, lam=0.0, dtau=0.01, loops=10, ring=True,
               desc="configs[3]: 4-D 256^4 fp32, time slabs over N GPUs, halos over NVLink inside the update kernel"),
    "c4s": dict(dims=(256, 256, 256, 64), real="f32", pot=0, m2=0.0, lam=0.0, dtau=0.01, loops=10, ring=True,
                desc="configs[3] at quarter size: 256^3 x 64 fp32, time slabs over N GPUs"),
}
BYTES_PER_UPDATE = {"f32": 8, "f64": 16}  # one read + one write of phi


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The timed
    region of the headline workload is tens of milliseconds, shorter than one `nvidia-smi -lms` period,
    so NVML is polled in-process every ~2 ms (nvidia-smi is the fallback)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nv, self.stop_flag = index, [], None, None, False
        self.sm, self.reasons, self.max_sm = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.nv is not None:
            self.stop_flag = True
            self.thread.join(1.0)
            return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def cpu_baseline(wl, seconds=12.0):
    """The oracle's OpenMP port (kind 'port': the reference itself is 1-D only and needs OpenCL)
    on a bounded sample of the same workload: same dims, dtau, seed; fewer tau-steps."""
    from oracle import oracle as O
    import numpy as np
    dims = wl["dims"]
    V = int(np.prod(dims))
    o = O.LatticeOracle(dims, real=O.F32 if wl["real"] == "f32" else O.F64, potential=wl["pot"], m2=wl["m2"],
                        lam=wl["lam"])
    cores = O.set_threads(len(os.sched_getaffinity(0)))
    o.step(wl["dtau"], 1, omp=True)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    o.step(wl["dtau"], 2, omp=True)
    per = (time.perf_counter() - t0) / 2
    n = max(2, min(2000, int(seconds / max(per, 1e-6))))
    t0 = time.perf_counter()
    o.step(wl["dtau"], n, omp=True)
    dt = time.perf_counter() - t0
    return {"value": V * n / dt, "unit": "site-updates/s", "cores": cores, "kind": "port",
            "sample": f"{n} tau-steps of the {'x'.join(map(str, dims))} lattice ({dt:.1f} s), oracle OpenMP port, "
                      f"{cores} threads"}, n, dt


def run_reference(args, wl, name):
    """--impl reference: the reference algorithm's CPU implementation on the host cores.  The
    reference's own kernel is 1-D / OpenCL-only, so for this lattice workload the oracle port
    (OpenMP, canonical chain+Jacobi semantics) stands in; each step = a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    import numpy as np
    dims = wl["dims"]
    V = int(np.prod(dims))
    o = O.LatticeOracle(dims, real=O.F32 if wl["real"] == "f32" else O.F64, potential=wl["pot"], m2=wl["m2"], lam=wl["lam"])
    cores = O.set_threads(len(os.sched_getaffinity(0)))  # torchrun sets OMP_NUM_THREADS=1: override
    o.step(wl["dtau"], 1, omp=True)
    t0 = time.perf_counter()
    o.step(wl["dtau"], 1, omp=True)
    per = time.perf_counter() - t0
    # size each step so the whole run stays within ~2 minutes
    total = args.steps + args.warmup
    sample = max(1, min(wl["loops"], int(100.0 / total / max(per, 1e-6))))
    for _ in range(args.warmup):
        o.step(wl["dtau"], sample, omp=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(wl["dtau"], sample, omp=True)
    dt = time.perf_counter() - t0
    val = V * sample * args.steps / dt
    line = {"impl": "reference", "metric": "lattice site-updates/s", "value": val, "unit": "site-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": wl["real"], "data": "synthetic",
            "config": {"workload": name, "desc": wl["desc"], "dims": list(dims), "dtau": wl["dtau"],
                       "tau_steps_per_step": sample},
            "cpu_baseline": {"value": val, "unit": "site-updates/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} of the workload's {wl['loops']} tau-steps per step, oracle OpenMP port, {cores} threads"},
            "e2e": {"value": val, "unit": "site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--math", default="fast", choices=["fast", "accurate"])
    ap.add_argument("--loops", type=int, default=0, help="tau-steps per bench step (default: workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    name = "c2" if args.workload == "auto" else args.workload
    wl = dict(WORKLOADS[name])
    if args.loops:
        wl["loops"] = args.loops
    if args.impl == "reference":
        return run_reference(args, wl, name)

    import numpy as np
    import torch
    import stochquant_b200 as sq

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dims = wl["dims"]
    V = int(np.prod(dims))
    # N > 1: each rank advances its own independent lattice of the same shape (north_star (4):
    # independent chains spread across GPUs, no data-path communication) -> weak scaling
    ring = bool(wl.get("ring"))
    sess = None
    if ring:
        # north_star (4): ONE lattice in time slabs, one per rank; total work fixed -> strong scaling
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from slab_common import split_slabs
        slab = split_slabs(dims[-1], world)[rank]
        # unique per launch (torchrun exports a run id): a crashed earlier run must not leave a segment we reuse
        run_id = "".join(ch for ch in os.environ.get("TORCHELASTIC_RUN_ID", "") if ch.isalnum())[:24]
        sess = sq.Session(f"bench{os.environ.get('MASTER_PORT', os.getpid())}{run_id}", rank, world)
        ctx = sq.Context(dims, real=wl["real"], math=args.math, potential=wl["pot"], m2=wl["m2"], lam=wl["lam"],
                         device=local, seed=1242608872, slab=slab)
        ctx.join(sess)
        args.no_e2e = True          # a 17 GB pinned host frame per step is not this workload's use
        args.no_cpu_baseline = True
    else:
        nch = int(wl.get("nchains", 1))
        if nch > 1:
            args.no_e2e = True      # frame_host moves one lattice; the batch of chains stays resident
        ctx = sq.Context(dims, real=wl["real"], math=args.math, potential=wl["pot"], m2=wl["m2"], lam=wl["lam"],
                         device=local, seed=1242608872 + rank, nchains=nch)
        if nch > 1:  # lambda grid x seeds, chains block-distributed over the ranks
            lams = np.linspace(0.0, 1.0, 64)
            for k in range(nch):
                g = rank * nch + k
                ctx.set_chain(k, 1242608872 + g // 64, wl["m2"], float(lams[g % 64]))
    nshare = (1 if ring else world) * int(wl.get("nchains", 1))   # lattices advanced by the job
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    loops, dtau = wl["loops"], wl["dtau"]
    for _ in range(args.warmup):
        ctx.step(dtau, loops)

    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = ctx.launch_count
    ms = []
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (outside the timed region)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        ctx.step_async(dtau, loops)
        ctx.sync()          # RNG-event replays are enqueued in here: they belong to the step
        e1.record(stream)
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = sum(ms)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = nshare * V * loops * args.steps / (total_ms * 1e-3)

    # ---- roofline: the update kernel alone, CUDA events around every launch -------------
    ctx.kernel_timing(True)
    ctx.step(dtau, loops)
    kms, kn = ctx.kernel_time()
    ctx.kernel_timing(False)
    peak, peak_src = peaks()
    bpu = BYTES_PER_UPDATE[wl["real"]]
    Vloc = ctx.vlocal * int(wl.get("nchains", 1))  # sites one launch of this rank's kernel updates
    units_per_launch = Vloc * loops / max(kn, 1)
    ach = units_per_launch * bpu / (kms / max(kn, 1) * 1e-3) / 1e9
    traffic, tnote = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            t = json.load(open(tp)).get(name)
            # ncu capture of one launch, scaled to the tau-steps one launch covers in this run
            traffic = t["bytes_per_launch"] * (units_per_launch / Vloc / t["tau_steps_per_launch"]) \
                if "lattice_" in t["kernel"] else t["bytes_per_launch"]
            tnote = t["source"]
        except Exception:
            traffic = None
    resident = len(dims) == 2 and wl["real"] == "f32" and dims[0] % 128 == 0 and dims[0] <= 1024 and dims[1] <= 8 * 148
    kname = "resident2d_kernel" if resident else ("lattice_march_kernel" if len(dims) >= 3 and wl["real"] == "f32"
                                                  else "lattice_step_kernel")
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": tnote,
                "kernel": kname, "launches_timed": kn,
                "tau_steps_per_launch": units_per_launch / Vloc,
                "avg_launch_us": 1e3 * kms / max(kn, 1), "bytes_per_site_update": bpu, "peak_source": peak_src,
                "note": ("on-chip resident kernel: the 4 MiB lattice is read/written once per launch, so `achieved` "
                         "(algorithmic bytes / time) measures instruction efficiency against the HBM roofline"
                         if resident else "streaming kernel: issue slots, L1 wavefronts and HBM are all at 50-60 % (DESIGN.md section 5)")}

    # ---- end to end through host buffers -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hin = torch.zeros(V, dtype=torch.float32 if wl["real"] == "f32" else torch.float64).pin_memory()
        hout = torch.empty_like(hin).pin_memory()
        hin.copy_(torch.from_numpy(ctx.download()))
        for _ in range(2):
            ctx.frame_host(hin.data_ptr(), hout.data_ptr(), dtau, loops, measure=True)
            hin.copy_(hout)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ctx.frame_host(hin.data_ptr(), hout.data_ptr(), dtau, loops, measure=True)
            hin, hout = hout, hin
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = V * (4 if wl["real"] == "f32" else 8)
        e2e = {"value": world * V * loops * args.steps / float(t.item()), "unit": "site-updates/s",
               "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes + 3 * 8 * dims[-1] + 64,
               "api": "sq_frame_host (pinned host field in, field + observables out)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_baseline(wl)

    if rank == 0:
        line = {"metric": "lattice site-updates/s", "value": value, "unit": "site-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if ring else "weak", "vs_baseline": None,
                "dtype": wl["real"], "data": "synthetic",
                "config": {"workload": name, "desc": wl["desc"], "dims": list(dims), "dtau": dtau,
                           "tau_steps_per_step": loops, "potential": wl["pot"], "math": args.math,
                           "seed": 1242608872, "l2": "flushed between timed steps (256 MB write)", "parallelism": (f"one lattice in {world} time slab(s), NVLink halo ring" if ring else
                                           f"{world} independent lattice(s), one per GPU"),
                           "slab": ctx.slab_stats() if ring else None},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()
    if sess is not None:
        sess.close()


if __name__ == "__main__":
    main()
