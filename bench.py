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

Beyond the contract's keys the default line carries
  `extras`  (N = 1) the other BASELINE configs measured in the same run: `accurate` (c2 with
            SQ_MATH_ACCURATE), `c2phi4_fast` (c2 with the phi^4 force), `c3` (64^4), `c5` (64 of the 512-chain share), `c1` (the drop-in
            ./tauhost.o on the reference's default command line, shortened);
  `ring`    (N > 1) configs[3]: ONE 256^4 lattice in N time slabs with the halos moving over NVLink
            inside the update kernel -- strong-scaling value, the same slab volume as a ring of one
            and as a plain single context on rank 0 (so the efficiency against the untaxed
            single-GPU rate is in the record), and `ring_parity`: before timing, a small lattice is
            advanced by the N-rank ring and by a single context and compared BITWISE (field + seed,
            with a forced RNG event); the run exits non-zero if they differ.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 1242608872
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
    # configs[0]: the reference's own default run through the drop-in executable (taumain.py:101-132)
    "c1": dict(dims=(200,), real="f64", pot=3, m2=0.0, lam=0.0, dtau=0.002, loops=1000, tauhost=True, frames=5000,
               desc="configs[0]: ./tauhost.o 200 0.02 0.002 5000 3 1.0 2 1 0 1000 0 <out> 40 (fp64, potID 3, "
                    "step-size controller on), stdout to /dev/null"),
}
BYTES_PER_UPDATE = {"f32": 8, "f64": 16}  # one read + one write of phi
L2_NOTE = "GPU arm: L2 flushed between timed steps (256 MB write); CPU arm: not applicable"
MATH_NOTE = {"fast": "GPU arm: SQ_MATH_FAST (SFU log2/sqrt/cos, bit-exact integer stream); CPU arm: libm",
             "accurate": "GPU arm: SQ_MATH_ACCURATE (CUDA logf/cosf/sqrtf with the reference's casts); CPU arm: libm"}


def make_config(name, wl, world, math):
    """The workload both arms run -- identical in `--impl ours` and `--impl reference` lines (what differs between
    the arms, e.g. the CPU arm's bounded sample, lives outside `config`)."""
    ring = bool(wl.get("ring"))
    return {"workload": name, "desc": wl["desc"], "dims": list(wl["dims"]), "dtau": wl["dtau"],
            "tau_steps_per_step": wl["loops"], "potential": wl["pot"], "math": MATH_NOTE[math], "seed": SEED,
            "l2": L2_NOTE,
            "parallelism": (f"one lattice in {world} time slab(s), NVLink halo ring" if ring else
                            f"{world} independent lattice(s), one per GPU (CPU arm: one lattice on the host cores)")}


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


# ------------------------------------------------------------------------------------------------
# CPU arms.  The oracle is test infrastructure; the only places it is executed outside tests/ are
# cpu_baseline() and run_reference() below, as the thing MEASURED BESIDE the product, never inside it.
def _lattice_oracle(wl):
    from oracle import oracle as O
    o = O.LatticeOracle(wl["dims"], real=O.F32 if wl["real"] == "f32" else O.F64, potential=wl["pot"], m2=wl["m2"],
                        lam=wl["lam"], seed=SEED)
    cores = O.set_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: override
    return o, cores


def cpu_baseline(wl, seconds=12.0):
    """The oracle's OpenMP port (kind 'port': the reference itself is 1-D only and needs OpenCL)
    on a bounded sample of the same workload: same dims, dtau, seed; fewer tau-steps."""
    import numpy as np
    V = int(np.prod(wl["dims"]))
    o, cores = _lattice_oracle(wl)
    o.step(wl["dtau"], 1, omp=True)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    o.step(wl["dtau"], 2, omp=True)
    per = (time.perf_counter() - t0) / 2
    n = max(2, min(2000, int(seconds / max(per, 1e-6))))
    t0 = time.perf_counter()
    o.step(wl["dtau"], n, omp=True)
    dt = time.perf_counter() - t0
    return {"value": V * n / dt, "unit": "site-updates/s", "cores": cores, "kind": "port",
            "sample": f"{n} tau-steps of the {'x'.join(map(str, wl['dims']))} lattice ({dt:.1f} s), oracle OpenMP port, "
                      f"{cores} threads"}


def c1_reference_frames(nframes, loops=1000):
    """configs[0] on the CPU with the REFERENCE'S OWN kernel source (oracle/_ref: tau_kernel.cl compiled by
    gcc, work-items as coroutines in gid order = what a CPU OpenCL runtime does with the single work-group of
    this run, tauhost.c:441-453) under the reference host's frame loop and step-size controller
    (tauhost.c:479-560).  One thread.  Returns (accepted site-updates, wall seconds, frames run)."""
    from oracle import oracle as O
    N, dt, dtau = 200, 0.02, 0.002
    f, om, r1 = O.host_init(N, dt, dtau)
    k = O.RefKernel(N, dt, dtau, 3, 1.0, f, om, r1)
    stab, runs = 0, 0
    t0 = time.perf_counter()
    for _ in range(nframes):
        f0, x0, xx00, om0 = k.f.copy(), k.x.copy(), k.xx0.copy(), k.omega.value
        k.launch(loops)
        if k.stable.value == 1:
            k.f[:], k.x[:], k.xx0[:] = k.newf, k.newx, k.newxx0
            if stab > 10:
                stab = 0
                k.deltaTau.value /= 0.95
            stab += 1
            runs += loops
            k.runs.value = runs
        else:
            k.f[:], k.x[:], k.xx0[:] = f0, x0, xx00
            k.omega.value = om0
            k.deltaTau.value *= 0.95
            stab = 0
            k.stable.value = 1
    return runs * N, time.perf_counter() - t0, nframes


def run_tauhost_c1(frames, loops=1000):
    """The drop-in executable on the reference's default command line (taumain.py:101-132), stdout to
    /dev/null, end file read back for the accepted tau-step count (`N` line, tauhost.c:577)."""
    exe = os.path.join(ROOT, "tauhost.o")
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "V0_2e_0-8.txt")
        args = [exe, "200", "0.02", "0.002", str(frames), "3", "1.0", "2", "1", "0", str(loops), "0", out, "40"]
        t0 = time.perf_counter()
        r = subprocess.run(args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError(f"tauhost.o exited {r.returncode}: {r.stderr[-400:]}")
        runs = 0
        for line in open(out):
            if line.rstrip().endswith("|N"):
                runs = int(line.split("|")[0])
    return {"frames": frames, "loops": loops, "wall_s": wall, "frames_per_s": frames / wall,
            "accepted_tau_steps": runs, "site_updates_per_s": 200 * runs / wall,
            "note": "process wall clock incl. CUDA context creation; rejected frames' partial steps are not counted"}


def run_reference(args, wl, name):
    """--impl reference: the reference algorithm's CPU implementation on the host cores.  For the lattice
    workloads the reference has no code (its kernel is 1-D / OpenCL-only): the oracle port (OpenMP, canonical
    chain+Jacobi semantics) stands in.  For c1 the reference's OWN kernel source runs (oracle/_ref).  Each step =
    a bounded sample of the workload's step (`cpu_baseline.sample`); `config` is the repo arm's."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    total = args.steps + args.warmup
    if wl.get("tauhost"):
        per_step = max(1, int(60.0 / total / 0.2))  # ~0.2 s per 1000-step frame, whole run ~1 minute
        for _ in range(args.warmup):
            c1_reference_frames(per_step, wl["loops"])
        upd, dt = 0, 0.0
        for _ in range(args.steps):
            u, t, _ = c1_reference_frames(per_step, wl["loops"])
            upd, dt = upd + u, dt + t
        val, cores, kind = upd / dt, 1, "reference"
        sample = (f"{per_step} frames x {wl['loops']} tau-steps per step from the cold start, the reference's own "
                  f"tau_kernel.cl compiled with gcc (oracle/_ref), 1 thread (one work-group)")
    else:
        V = int(np.prod(wl["dims"]))
        o, cores = _lattice_oracle(wl)
        o.step(wl["dtau"], 1, omp=True)
        t0 = time.perf_counter()
        o.step(wl["dtau"], 1, omp=True)
        per = time.perf_counter() - t0
        # size each step so the whole run stays within ~2 minutes
        n = max(1, min(wl["loops"], int(100.0 / total / max(per, 1e-6))))
        for _ in range(args.warmup):
            o.step(wl["dtau"], n, omp=True)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            o.step(wl["dtau"], n, omp=True)
        dt = time.perf_counter() - t0
        val, kind = V * n * args.steps / dt, "port"
        sample = f"{n} of the workload's {wl['loops']} tau-steps per step, oracle OpenMP port, {cores} threads"
    line = {"impl": "reference", "metric": "lattice site-updates/s", "value": val, "unit": "site-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "strong" if wl.get("ring") else "weak", "vs_baseline": None,
            "dtype": wl["real"], "data": "synthetic",
            "config": make_config(name, wl, args.gpus, "accurate" if wl.get("tauhost") else args.math),
            "cpu_baseline": {"value": val, "unit": "site-updates/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Timer:
    """Frames of one context timed with CUDA events on the library's stream (step_async + sync inside)."""

    def __init__(self, torch, local, flush):
        self.torch, self.local, self.flush = torch, local, flush

    def frames(self, ctx, dtau, loops, steps, warmup):
        torch = self.torch
        stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", self.local))
        for _ in range(warmup):
            ctx.step(dtau, loops)
        ms = []
        for _ in range(steps):
            self.flush.fill_(1)  # L2 flush between timed iterations (outside the timed region)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            ctx.step_async(dtau, loops)
            ctx.sync()          # RNG-event replays are enqueued in here: they belong to the step
            e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        return ms

    def kernel(self, ctx, dtau, loops):
        """average duration of the dominant (update) kernel: events around every launch"""
        ctx.kernel_timing(True)
        ctx.step(dtau, loops)
        kms, kn = ctx.kernel_time()
        ctx.kernel_timing(False)
        return kms, kn


def kernel_name(dims, real):
    """the dominant kernel of a workload (mirrors the library's dispatch, sq_api.cu: init_lattice)"""
    resident = False
    if len(dims) == 2 and real == "f32" and dims[0] % 128 == 0 and dims[0] <= 1024 and dims[1] >= 2:
        nb = min(148, dims[1] // 2)
        rows = -(-dims[1] // nb)
        resident = (dims[0] % 256 == 0 and rows * (dims[0] // 8) <= 896) or rows * (dims[0] // 4) <= 896
    if resident:
        return "rowres_kernel", True
    if len(dims) >= 3 and real == "f32":
        # bulk-copy staged tiles for event-free steps (the timed launches), the marching kernel where a step carries replay entries
        return ("lattice_rows_kernel" if os.environ.get("SQ_ROWS") == "1" else "lattice_tile_kernel"), False
    return "lattice_step_kernel", False


def roofline_of(name, wl, ctx, kms, kn, loops):
    peak, peak_src = peaks()
    bpu = BYTES_PER_UPDATE[wl["real"]]
    Vloc = ctx.vlocal * int(wl.get("nchains", 1))  # sites one launch of this rank's kernel updates
    units_per_launch = Vloc * loops / max(kn, 1)
    ach = units_per_launch * bpu / (kms / max(kn, 1) * 1e-3) / 1e9
    traffic, tnote = None, None
    kname, resident = kernel_name(wl["dims"], wl["real"])
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            t = json.load(open(tp)).get(name)
            # ncu capture of one launch of this workload's kernel covering the same number of tau-steps (a few replayed
            # steps -- launches of the marching kernel after an RNG event -- shift the average by a few per cent)
            if t["kernel"].startswith(kname) and abs(units_per_launch / Vloc - t["tau_steps_per_launch"]) <= 0.05 * t["tau_steps_per_launch"]:
                traffic = t["bytes_per_launch"]
                tnote = t["source"]
        except Exception:
            traffic = None
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": tnote, "kernel": kname, "launches_timed": kn,
            "tau_steps_per_launch": units_per_launch / Vloc,
            "avg_launch_us": 1e3 * kms / max(kn, 1), "bytes_per_site_update": bpu, "peak_source": peak_src,
            "note": ("on-chip resident kernel: the 4 MiB lattice is read/written once per launch, so `achieved` "
                     "(algorithmic bytes / time) measures instruction efficiency against the HBM roofline"
                     if resident else "streaming kernel: algorithmic 8 B per site-update against measured HBM copy bandwidth")}


def make_context(sq, np, wl, math, local, rank, nchains=None, slab=(0, 0), dims=None):
    nch = int(wl.get("nchains", 1)) if nchains is None else nchains
    ctx = sq.Context(dims or wl["dims"], real=wl["real"], math=math, potential=wl["pot"], m2=wl["m2"], lam=wl["lam"],
                     device=local, seed=SEED + (rank if not wl.get("ring") and slab == (0, 0) else 0), nchains=nch,
                     slab=slab)
    if nch > 1:  # lambda grid x seeds, chains block-distributed over the ranks
        lams = np.linspace(0.0, 1.0, 64)
        for k in range(nch):
            g = rank * nch + k
            ctx.set_chain(k, SEED + g // 64, wl["m2"], float(lams[g % 64]))
    return ctx


def session_name(dist, tag):
    """Unique per launch AND per ring: rank 0 draws a nonce, everybody gets it (a crashed earlier run must not
    leave a segment this ring would attach to)."""
    nonce = [f"{os.getpid():x}{time.time_ns() & 0xFFFFFFFFFF:x}"]
    if dist is not None:
        dist.broadcast_object_list(nonce, src=0)
    return f"b{tag}{nonce[0]}"


# seed whose chain meets an inf-retry (tau_kernel.cl:282) at gid 4099 of the first step -- inside rank 0's slab
# of the parity lattice, so every other rank has to apply the agreed correction (tests/helpers.py derives it)
RING_PARITY_SEED = 244480037558186


def ring_parity(sq, np, torch, dist, world, rank, local):
    """Small lattice 32x16x8x(8N): the N-rank ring over NVLink vs one context on rank 0, compared bitwise."""
    from stochquant_b200.slabs import split_slabs
    dims = (32, 16, 8, 8 * world)
    V = int(np.prod(dims))
    vs = V // dims[-1]
    rng = np.random.default_rng(11)
    phi0 = (rng.normal(size=V) * 0.5).astype(np.float32)
    t0, nt = split_slabs(dims[-1], world)[rank]
    sess = sq.Session(session_name(dist, "p"), rank, world)
    ctx = sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, device=local, seed=RING_PARITY_SEED,
                     slab=(t0, nt))
    ctx.upload(phi0[t0 * vs:(t0 + nt) * vs])
    ctx.join(sess)
    for n in (5, 20):
        ctx.step(0.01, n)
    m = ctx.measure()
    mine = torch.from_numpy(ctx.download().copy()).cuda()
    seed = torch.tensor([int(m["seed"]) & (2**63 - 1), int(m["nevents"])], dtype=torch.int64, device="cuda")
    ctx.close()
    sess.close()
    parts = [torch.empty(split_slabs(dims[-1], world)[r][1] * vs, dtype=torch.float32, device="cuda") for r in range(world)]
    seeds = [torch.empty(2, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(parts, mine)
    dist.all_gather(seeds, seed)
    verdict = None
    if rank == 0:
        one = sq.Context(dims, real="f32", math="fast", potential=4, m2=0.25, lam=0.5, device=local, seed=RING_PARITY_SEED)
        one.upload(phi0)
        for n in (5, 20):
            one.step(0.01, n)
        m1 = one.measure()
        ref = one.download()
        one.close()
        got = torch.cat(parts).cpu().numpy()
        same = bool(np.array_equal(got.view(np.uint32), ref.view(np.uint32)))
        same &= all(int(s[0].item()) == (int(m1["seed"]) & (2**63 - 1)) for s in seeds)
        same &= int(m1["nevents"]) >= 1 and all(int(s[1].item()) >= 1 for s in seeds)
        verdict = "bit-identical" if same else "FAILED"
    return verdict


NVLINK_STATUS = {"source": None, "why": None}


def nvlink_bytes(local):
    """NVLink data bytes (tx, rx) of this rank's GPU since driver load, summed over its links: NVML field values
    (KiB counters), else `nvidia-smi nvlink -gt d`; None where neither reports them (NVLINK_STATUS says why)."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
    why = []
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ids = [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX]
        vals = pynvml.nvmlDeviceGetFieldValues(h, ids)
        if all(v.nvmlReturn == 0 for v in vals):
            NVLINK_STATUS.update(source="nvml field values", why=None)
            return [int(v.value.ullVal) * 1024 for v in vals]
        why.append("nvml field values: nvmlReturn %s" % [int(v.nvmlReturn) for v in vals])
    except Exception as e:  # noqa: BLE001
        why.append("nvml: %s: %s" % (type(e).__name__, e))
    try:
        import re
        import subprocess
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(phys)], capture_output=True, text=True, timeout=20).stdout
        tx = [int(m) for m in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt)]
        rx = [int(m) for m in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt)]
        if tx and rx:
            NVLINK_STATUS.update(source="nvidia-smi nvlink -gt d", why=None)
            return [sum(tx) * 1024, sum(rx) * 1024]
        why.append("nvidia-smi nvlink -gt d: no counters in %r" % txt[:120])
    except Exception as e:  # noqa: BLE001
        why.append("nvidia-smi: %s: %s" % (type(e).__name__, e))
    NVLINK_STATUS.update(source=None, why="; ".join(why))
    return None


def ring_block(sq, np, torch, dist, timer, world, rank, local, args):
    """configs[3] inside the multi-GPU line: the 256^4 ring on all ranks, then -- on rank 0 alone, the other GPUs
    idle -- the same slab volume as a ring of one (finder + halo protocol, no NVLink) and as a plain context."""
    from stochquant_b200.slabs import split_slabs
    wl = dict(WORKLOADS["c4"])
    dims, loops, dtau = wl["dims"], wl["loops"], wl["dtau"]
    V = int(np.prod(dims))
    parity = ring_parity(sq, np, torch, dist, world, rank, local)
    t0, nt = split_slabs(dims[-1], world)[rank]
    sess = sq.Session(session_name(dist, "r"), rank, world)
    ctx = make_context(sq, np, wl, args.math, local, rank, slab=(t0, nt))
    ctx.join(sess)
    steps, warm = max(3, min(args.steps, 6)), 2

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warm):
        ctx.step(dtau, loops)
    # (the counters are read OUTSIDE the barrier-bracketed region: reading them takes a rank-dependent while, and in a ring
    # one late rank is waited for by all the others inside their first timed step)
    nv0 = nvlink_bytes(local) if rank == 0 else None
    barrier()
    ms = timer.frames(ctx, dtau, loops, steps, 0)
    barrier()
    nv1 = nvlink_bytes(local) if rank == 0 else None
    t = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    kms, kn = timer.kernel(ctx, dtau, loops)
    stats = ctx.slab_stats()
    nsteps_total = (warm + steps + 1) * loops
    vloc = ctx.vlocal
    ctx.close()
    sess.close()
    barrier()
    out = None
    if rank == 0:
        value = V * loops * steps / (total_ms * 1e-3)
        peak, _ = peaks()
        kus = 1e3 * kms / max(kn, 1)
        out = {"workload": "c4", "desc": wl["desc"], "dims": list(dims), "scaling": "strong", "value": value,
               "unit": "site-updates/s", "ms_per_step": total_ms / steps, "steps": steps, "tau_steps_per_step": loops,
               "update_kernel_us": kus, "roofline_frac": vloc * 8 / (kus * 1e-6) / 1e9 / peak,
               "halo_bytes_per_tau_step_per_rank": {"sent": 2 * (V // dims[-1]) * 4, "received": 2 * (V // dims[-1]) * 4},
               # NVML's NVLink data counters of rank 0's GPU around the timed frames (includes the L2 flushes' nothing: the
               # flush is local) -- the halo slices are the only peer traffic of the run
               "nvlink_bytes_per_tau_step_rank0": ({"tx": (nv1[0] - nv0[0]) / (steps * loops), "rx": (nv1[1] - nv0[1]) / (steps * loops),
                                                    "source": NVLINK_STATUS["source"]} if nv0 and nv1 else None),
               "nvlink_counters_unavailable": (None if nv0 and nv1 else NVLINK_STATUS["why"]),
               "finder_scans_per_tau_step": stats["finder_scans"] / nsteps_total,
               "agree_rounds_per_tau_step": stats["agree_rounds"] / nsteps_total,
               "ring_parity": parity,
               "ring_parity_what": "32x16x8x(8N) lattice, 25 tau-steps, forced RNG event in rank 0's slab: N-rank ring vs "
                                   "one context, field and seed compared bitwise"}
        # the same slab volume on ONE GPU: ring of one (pays the finder), plain context (does not)
        sdims = tuple(dims[:-1]) + (nt,)
        wls = dict(wl, dims=sdims)
        s1 = sq.Session(f"bs{os.getpid():x}{time.time_ns() & 0xFFFFFFFF:x}", 0, 1)
        c1 = sq.Context(sdims, real="f32", math=args.math, device=local, seed=SEED, slab=(0, nt))
        c1.join(s1)
        m1 = timer.frames(c1, dtau, loops, 3, 2)
        c1.close()
        s1.close()
        cp = sq.Context(sdims, real="f32", math=args.math, device=local, seed=SEED)
        mp = timer.frames(cp, dtau, loops, 3, 2)
        kp, knp = timer.kernel(cp, dtau, loops)
        cp.close()
        vs_ = int(np.prod(sdims))
        r1 = vs_ * loops * 3 / (sum(m1) * 1e-3)
        rp = vs_ * loops * 3 / (sum(mp) * 1e-3)
        out["single_gpu_same_slab"] = {"dims": list(sdims), "ring_of_one": r1, "plain_context": rp,
                                       "plain_update_kernel_us": 1e3 * kp / max(knp, 1),
                                       "finder_tax": 1.0 - r1 / rp}
        out["efficiency_vs_plain_per_gpu"] = value / (world * rp)
        out["efficiency_vs_ring_of_one"] = value / (world * r1)
        out["limiter"] = ("event finder (integer scan of every draw ahead of the update, shared by the ranks) + the "
                          "boundary slices' wait for the neighbour's flag; halo bytes are ~3 % of a rank's HBM traffic")
        del wls
    barrier()
    return out, parity


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
    ap.add_argument("--no-extras", action="store_true", help="skip the `extras` / `ring` blocks of the default line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    auto = args.workload == "auto"
    name = "c2" if auto else args.workload
    wl = dict(WORKLOADS[name])
    if args.loops:
        wl["loops"] = args.loops
    if args.impl == "reference":
        return run_reference(args, wl, name)

    import numpy as np
    import torch
    import stochquant_b200 as sq
    from stochquant_b200.slabs import split_slabs

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

    if wl.get("tauhost"):   # configs[0]: a process, not a context
        if rank == 0:
            run_tauhost_c1(64)  # warm-up: CUDA context creation, lazy module load, page cache
            r = run_tauhost_c1(wl["frames"], wl["loops"])
            upd, wall, nfr = c1_reference_frames(100, wl["loops"])
            line = {"metric": "lattice site-updates/s", "value": r["site_updates_per_s"], "unit": "site-updates/s",
                    "n_gpus": 1, "steps": wl["frames"], "warmup": 64, "ms_per_step": 1e3 * r["wall_s"] / wl["frames"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                    "config": make_config(name, wl, 1, "accurate"), "tauhost": r,
                    "roofline": None, "roofline_note": "4.8 KB of state in one CTA: latency-bound, no roofline fraction (SURVEY.md 8(d) C1)",
                    "cpu_baseline": {"value": upd / wall, "unit": "site-updates/s", "cores": 1, "kind": "reference",
                                     "sample": f"{nfr} frames x {wl['loops']} tau-steps from the cold start ({wall:.1f} s), the "
                                               "reference's own tau_kernel.cl compiled with gcc (oracle/_ref), 1 thread"},
                    "e2e": {"value": r["site_updates_per_s"], "unit": "site-updates/s", "h2d_bytes_per_step": 0,
                            "d2h_bytes_per_step": 8 * 200 + 16,
                            "api": "./tauhost.o process: argv in, stdout frame stream + end file out"},
                    "gpu_launches": wl["frames"] + 1}
            print(json.dumps(line), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    dims = wl["dims"]
    V = int(np.prod(dims))
    # N > 1: each rank advances its own independent lattice of the same shape (north_star (4):
    # independent chains spread across GPUs, no data-path communication) -> weak scaling
    ring = bool(wl.get("ring"))
    sess = None
    if ring:
        # north_star (4): ONE lattice in time slabs, one per rank; total work fixed -> strong scaling
        slab = split_slabs(dims[-1], world)[rank]
        sess = sq.Session(session_name(dist, "w"), rank, world)
        ctx = make_context(sq, np, wl, args.math, local, rank, slab=slab)
        ctx.join(sess)
        args.no_e2e = True          # a 17 GB pinned host frame per step is not this workload's use
        args.no_cpu_baseline = True
    else:
        if int(wl.get("nchains", 1)) > 1:
            args.no_e2e = True      # frame_host moves one lattice; the batch of chains stays resident
        ctx = make_context(sq, np, wl, args.math, local, rank)
    nshare = (1 if ring else world) * int(wl.get("nchains", 1))   # lattices advanced by the job
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    timer = Timer(torch, local, flush)

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
    ms = timer.frames(ctx, dtau, loops, args.steps, 0)
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
    kms, kn = timer.kernel(ctx, dtau, loops)
    roofline = roofline_of(name, wl, ctx, kms, kn, loops)
    nevents = int(ctx.measure()["nevents"])

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
    slab_stats = ctx.slab_stats() if ring else None
    ctx.close()
    if sess is not None:
        sess.close()

    # ---- the other BASELINE configs, measured in the same run --------------------------------
    extras, ringblk, parity = None, None, None
    if auto and not args.no_extras:
        if world == 1:
            extras = {}
            k = max(3, min(args.steps, 5))
            # c2 with the reference's casts and CUDA's logf/cosf/sqrtf (SQ_MATH_ACCURATE)
            # ... and with the phi^4 force (BASELINE configs[1] says "phi^4": SURVEY 8(d) C2 fixes potID 0 with this as its variant)
            for wname, math, nch in (("c2", "accurate", None), ("c2phi4", "fast", None), ("c3", "fast", None), ("c3", "accurate", None),
                                     ("c5", "fast", 64)):
                w = dict(WORKLOADS[wname])
                if math == args.math and wname == name:
                    continue
                c = make_context(sq, np, w, math, local, 0, nchains=nch)
                m_ = timer.frames(c, w["dtau"], w["loops"], k, 2)
                km, kn_ = timer.kernel(c, w["dtau"], w["loops"])
                nchn = nch or int(w.get("nchains", 1))
                rf = roofline_of(wname, dict(w, nchains=nchn), c, km, kn_, w["loops"])
                c.close()
                Vw = int(np.prod(w["dims"])) * nchn
                extras[f"{wname}_{math}"] = {"value": Vw * w["loops"] * k / (sum(m_) * 1e-3), "unit": "site-updates/s",
                                             "ms_per_step": sum(m_) / k, "steps": k, "dims": list(w["dims"]), "nchains": nchn,
                                             "tau_steps_per_step": w["loops"], "math": math, "frac": rf["frac"],
                                             "kernel": rf["kernel"], "avg_launch_us": rf["avg_launch_us"]}
            extras["accurate"] = {"value": extras["c2_accurate"]["value"], "frac": extras["c2_accurate"]["frac"]} \
                if "c2_accurate" in extras else None
            try:
                run_tauhost_c1(64)
                extras["c1"] = run_tauhost_c1(500)
                extras["c1"]["command"] = "./tauhost.o 200 0.02 0.002 500 3 1.0 2 1 0 1000 0 <out> 40 > /dev/null (configs[0] at a tenth of its 5000 frames)"
            except Exception as e:  # noqa: BLE001
                extras["c1"] = {"error": str(e)[:300]}
        else:
            ringblk, parity = ring_block(sq, np, torch, dist, timer, world, rank, local, args)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(wl)

    if rank == 0:
        cfg = make_config(name, wl, world, args.math)
        line = {"metric": "lattice site-updates/s", "value": value, "unit": "site-updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
                "higher_is_better": True, "scaling": "strong" if ring else "weak", "vs_baseline": None,
                "dtype": wl["real"], "data": "synthetic", "config": cfg, "slab": slab_stats,
                "rng_events_replayed": nevents,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if extras is not None:
            line["extras"] = extras
        if ringblk is not None:
            line["ring"] = ringblk
            line["ring_parity"] = parity
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity == "FAILED":
        sys.exit(4)


if __name__ == "__main__":
    main()
