#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native SocialMediaModel hot path.

Contract (one JSON line on stdout from rank 0):
    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference algorithm on host cores

Default workload = BASELINE.json configs[1] (C2): ABMsolveensemble, 1 024 realizations of the
default-parameter ABM (n=250, L=4, M=2, NT=200, 4 save points, KDE on the 89x89 grid) PER GPU
(weak scaling: rank r runs realizations r*1024+1 ..; the ensemble-mean density is all-reduced over
NCCL).  A "step" is one full pass: 1 024 x 200 time steps + KDE + ensemble sums.
metric = ABM agent-steps/s (whole job).  `value`: results stay in HBM.  `e2e`: the same pass
through the host-buffer C-ABI call the Julia shim makes (smm_abm_ensemble), D2H of us/zs/ys
included, with its measured floor (PCIe and host-write rates of this box).
"workloads" carries, at every N: the end-to-end variants (pageable caller array, Float32 densities,
means only), configs[1] as written (1 024 realizations over ALL GPUs of the job; at N>1 also through
ONE call from one process, smm_ctx_create_multi) and the PDE half of the metric (C4: 64 instances
per GPU at 353x353, sharded by rank, means all-reduced); at N=1 also C1, C5 and C3.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2 = dict(R=1024, NT=200, n=250, L=4, M=2, savepoints=4, kde_sigma=0.1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """start of the timed region: only samples taken after this are summarised"""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        rows = [r for t, r in self.rows if t >= t_mark] or [r for t, r in self.rows[-3:]]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return float(x)
    import torch.distributed as dist
    t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------
# C2: ABM ensemble (default)
# ------------------------------------------------------------------------------------------
# DRAM traffic per launch of the dominant kernels, from `ncu --set full` (dram__bytes_read.sum + dram__bytes_write.sum;
# the .csv pages are under profiles/): kernel -> (bytes per launch at the benchmarked size, source)
NCU_TRAFFIC = {
    "abm_dense": (3.61e6, "profiles/r2c_abm_dense_full_raw.csv (1 024 realizations x 200 steps in chunks of 20 steps; the state lives in shared memory, "
                          "4.3 KB per realization and chunk boundary are handed over through L2)"),
    "kde": (9.93e8, "profiles/r2c_kde_full_raw.csv (980 MB written + 12.6 MB read; algorithmic output 1.04 GB)"),
    "pop_step": (1.03e9, "profiles/r1_pop_step_full_raw_v2.csv (n = 1e7; positions are L2-resident)"),
    # per launch at 16 instances 353 x 353 (profiles/r2_pde_final_full_raw.csv); scaled by B / 16 in bench_pde.py
    "pde_stage2_per16": (3.28e8, "profiles/r2_pde_final_full_raw.csv: 210 MB read + 117 MB written per launch at 16 x 353^2 (algorithmic 303 MB)"),
}


class AbmEnsemble:
    """Device buffers + the two timed calls of one ensemble shard of R realizations on this rank's GPU."""

    def __init__(self, ctx, stream, R, first, cfg=C2):
        import ctypes as C

        import torch

        import socialmediamodel_b200 as smm
        from socialmediamodel_b200 import _lib
        self.C, self._lib, self.ctx, self.stream, self.R, self.first, self.cfg = C, _lib, ctx, stream, R, first, cfg
        NT, n, L, M, sp = cfg["NT"], cfg["n"], cfg["L"], cfg["M"], cfg["savepoints"]
        self.p, self.q = smm.ABMconstruct(), smm.parameters(n=n, L=L)
        self.Nx, self.Ny = self.p.X.shape
        self.savesteps = smm.savetimes_abm(NT, sp)
        self.per_real = self.Nx * self.Ny * 2 * L * sp
        self.us_dev = torch.empty(R * self.per_real, dtype=torch.float32, device="cuda")
        self.sum_us = torch.zeros(self.per_real, dtype=torch.float64, device="cuda")
        self.sum_zs = torch.zeros(2 * 2 * sp, dtype=torch.float64, device="cuda")
        self.sum_ys = torch.zeros(2 * L * sp, dtype=torch.float64, device="cuda")
        self.numbers = torch.zeros(M * L * (NT + 1), dtype=torch.int64, device="cuda")
        self.spar, self.grid = _lib.make_params(self.q, self.p), _lib.make_grid(self.p)

    def device_pass(self, world):
        C, cfg, ctx = self.C, self.cfg, self.ctx
        rc = ctx.lib.smm_abm_ensemble_dev(ctx.h, C.byref(self.spar), 0, cfg["NT"], self.R, self.first, 0, None, None, cfg["savepoints"],
                                          self._lib.ptr(self.savesteps), cfg["kde_sigma"], C.byref(self.grid),
                                          C.c_void_p(self.us_dev.data_ptr()), C.c_void_p(self.sum_us.data_ptr()),
                                          C.c_void_p(self.sum_zs.data_ptr()), C.c_void_p(self.sum_ys.data_ptr()),
                                          C.c_void_p(self.numbers.data_ptr()))
        ctx.check(rc, "smm_abm_ensemble_dev")
        if world > 1:   # ensemble mean over all GPUs (plotting.jl:361-363): the path's only exchange
            import torch.distributed as dist
            dist.all_reduce(self.sum_us); dist.all_reduce(self.sum_zs); dist.all_reduce(self.sum_ys)

    def time_device(self, steps, warmup, world, flush):
        """K passes, CUDA events on the launching stream, L2 flushed between passes; max over ranks."""
        import torch
        ctx, stream = self.ctx, self.stream
        for _ in range(warmup):
            self.device_pass(world); flush.fill_(1)
        barrier(world)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        kms = {"abm": [], "kde": [], "reduce": []}
        l0 = ctx.launches
        barrier(world)
        t0 = time.perf_counter()
        for k in range(steps):
            ev[k][0].record(stream)
            self.device_pass(world)
            ev[k][1].record(stream)
            flush.fill_(k & 1)                       # L2 flush between timed passes (outside the event pair)
            ctx.sync()                               # collects the per-kernel CUDA-event times of this pass
            kms["abm"].append(ctx.kernel_ms(0)); kms["kde"].append(ctx.kernel_ms(1)); kms["reduce"].append(ctx.kernel_ms(2))
        barrier(world)
        wall = time.perf_counter() - t0
        dev_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), world)
        return dict(dev_ms=dev_ms, wall_s=wall, launches=ctx.launches - l0, kms={k: float(np.mean(v)) for k, v in kms.items()})

    def time_e2e(self, steps, world, mode="f64", pinned=True, ctx=None, R=None, first=None):
        """The host-buffer C-ABI call the Julia shim makes (smm_abm_ensemble), wall clock around the blocking call, max over
        ranks.  mode: "f64" = the reference's return type (us Float64), "f32" = smm_abm_ensemble_f32, "means" = us NULL
        (ensemble sums only).  pinned=False: the caller's `us` is ordinary pageable memory, as the Julia shim allocates it."""
        C, _lib, cfg = self.C, self._lib, self.cfg
        ctx = ctx or self.ctx
        R = R or self.R
        first = self.first if first is None else first
        NT, L, sp = cfg["NT"], cfg["L"], cfg["savepoints"]
        shape = (self.Nx, self.Ny, 2, L, sp, R)
        dt = np.float32 if mode == "f32" else np.float64
        out_us = None
        if mode != "means":
            out_us = ctx.pinned_empty(shape, dt) if pinned else np.zeros(shape, dt, order="F")   # zeros(): what the shim does
        zs = np.empty((2, 2, sp, R), order="F"); ys = np.empty((2, L, sp, R), order="F")
        s_us = np.empty((self.Nx, self.Ny, 2, L, sp), order="F"); s_zs = np.empty((2, 2, sp), order="F"); s_ys = np.empty((2, L, sp), order="F")
        ns = _lib.smm_noise(); ns.mode = 0; ns.seed = 0
        tail = (None, None, None, None, None, None, None, C.byref(ns), sp, _lib.ptr(self.savesteps), cfg["kde_sigma"],
                C.byref(self.grid), _lib.ptr(out_us), _lib.ptr(zs), _lib.ptr(ys), _lib.ptr(s_us), _lib.ptr(s_zs), _lib.ptr(s_ys), None)

        def one():
            if mode == "f32":
                rc = ctx.lib.smm_abm_ensemble_f32(ctx.h, C.byref(self.spar), NT, R, first, *tail)
            else:
                rc = ctx.lib.smm_abm_ensemble(ctx.h, C.byref(self.spar), 0, NT, R, first, *tail)
            ctx.check(rc, "smm_abm_ensemble")

        one()
        barrier(world)
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        barrier(world)
        secs = max_over_ranks(time.perf_counter() - t0, world)
        small = zs.nbytes + ys.nbytes + s_us.nbytes + s_zs.nbytes + s_ys.nbytes
        # the library widens FP32 -> Float64 on the host only for chunked shards (>= 512 realizations per GPU) and only
        # when the GPU's share of the host cores is >= 12 threads (api_abm.cu); otherwise Float64 crosses PCIe
        hw = os.environ.get("SMM_HOST_WIDEN")
        multi = ctx.gpu_count > 1
        share = ctx.gpu_count if multi else int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        threads = min(16, (os.cpu_count() or 1) // max(1, share))
        widen = mode == "f64" and R // ctx.gpu_count >= 512 and (hw == "1" or (hw != "0" and threads >= 12))
        us_bytes = 0 if out_us is None else out_us.nbytes
        d2h = (us_bytes // 2 if widen else us_bytes) + small
        mass = float(s_us.sum() * self.p.dV / (R * sp))
        if out_us is not None and pinned:
            ctx.free_pinned(out_us)
        return dict(ms_per_step=secs / steps * 1e3, steps=steps, d2h_bytes_per_step=d2h,
                    h2d_bytes_per_step=C.sizeof(_lib.smm_params) + C.sizeof(_lib.smm_grid) + self.savesteps.nbytes,
                    host_bytes_written_per_step=us_bytes + small, host_widen=bool(widen), mass_check=mass,
                    caller_array="pinned (smm_host_alloc)" if pinned else "pageable zeros(), as the Julia shim allocates it")


def e2e_floor(ctx, host_bytes_f64, widen):
    """Floor of the full-`us` call on this box: the densities must cross PCIe (as FP32 when the host widens, else Float64)
    and be written into the caller's Float64 array; both rates are measured here (smm_microbench), the two overlap."""
    d2h = ctx.microbench("d2h", 256 << 20)
    wid = ctx.microbench("widen", 256 << 20)
    pcie_ms = (host_bytes_f64 / 2 if widen else host_bytes_f64) / (d2h * 1e9) * 1e3
    host_ms = host_bytes_f64 / (wid * 1e9) * 1e3 if widen else 0.0
    return dict(floor_ms=max(pcie_ms, host_ms), pcie_d2h_gbs=d2h, host_widen_gbs=wid, pcie_ms=pcie_ms, host_write_ms=host_ms)


def run_abm_ensemble(args, rank, world, local):
    import torch

    import socialmediamodel_b200 as smm

    cfg = dict(C2)
    if args.realizations:
        cfg["R"] = args.realizations
    R, NT, n, L, M, sp = cfg["R"], cfg["NT"], cfg["n"], cfg["L"], cfg["M"], cfg["savepoints"]
    ctx = smm.Context(local)
    stream = torch.cuda.Stream()             # a real stream (the legacy default stream cannot be graph-captured)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    ens = AbmEnsemble(ctx, stream, R, rank * R + 1, cfg)               # seeds k = 1..N (ensemble.jl:25)
    sampler = ClockSampler(local); sampler.start()   # started before the warm-up so it is sampling by the timed region
    for _ in range(args.warmup):
        ens.device_pass(world); flush.fill_(1)
    barrier(world)
    sampler.mark()
    t = ens.time_device(args.steps, 0, world, flush)
    clocks = sampler.stop()
    dev_ms, kms, launches = t["dev_ms"], t["kms"], t["launches"]
    value = world * R * NT * n * args.steps / (dev_ms * 1e-3)

    # ---- e2e: host-buffer C-ABI call (what the Julia shim's ABMsolveensemble makes) ------------
    e2e_steps = max(1, min(args.steps, 5))
    e = ens.time_e2e(e2e_steps, world, mode="f64", pinned=True)
    e2e_value = world * R * NT * n / (e["ms_per_step"] * 1e-3)
    floor = e2e_floor(ctx, R * ens.per_real * 8, e["host_widen"]) if rank == 0 else None

    peaks = measured_peaks()
    abm_ms = kms["abm"]
    pairs = R * NT * n * n                                              # ordered pairs per launch
    sym = os.environ.get("SMM_ABM_SYM", "1") != "0"
    mufu_per_pair = 1 if sym else 2    # sqrt + ex2 per evaluated pair; the symmetric kernel evaluates unordered pairs
    mufu_meas = ctx.microbench("mufu")                                   # G MUFU/s, measured on this GPU right now
    mufu_nominal = 148 * 16 * peaks["sm_max_mhz"] * 1e6 / 1e9
    mufu_peak = mufu_meas / mufu_per_pair                                # G ordered pairs/s
    roofline = dict(kernel="abm_dense_kernel<float,sym>" if sym else "abm_dense_kernel<float>", bound="mufu",
                    achieved=pairs / (abm_ms * 1e-3) / 1e9, peak=mufu_peak, unit="Gpair/s",
                    frac=pairs / (abm_ms * 1e-3) / 1e9 / mufu_peak, traffic=NCU_TRAFFIC["abm_dense"][0],
                    traffic_source=NCU_TRAFFIC["abm_dense"][1],
                    note="state lives in shared memory; not HBM-bound (SURVEY 8d). Algorithmic work: 2 MUFU (sqrt, ex2) per "
                         "UNORDERED pair = 1 per ordered pair (w_ij = w_ji); peak = MUFU issue rate MEASURED by smm_microbench "
                         f"(ex2.approx chains, {mufu_meas:.0f} G/s; nominal 148 SM x 16/clk x {peaks['sm_max_mhz']:.0f} MHz = "
                         f"{mufu_nominal:.0f} G/s)",
                    mufu_measured_gops=mufu_meas, mufu_nominal_gops=mufu_nominal,
                    kernel_ms=abm_ms, share_of_step=abm_ms * args.steps / dev_ms if world == 1 else None)
    kde_ms = kms["kde"]
    kde_bytes = R * ens.per_real * 4
    roofline_kde = dict(kernel="kde_kernel<float>", bound="hbm", achieved=kde_bytes / (kde_ms * 1e-3) / 1e9,
                        peak=peaks["hbm_gbs"], unit="GB/s", frac=kde_bytes / (kde_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        traffic=NCU_TRAFFIC["kde"][0], kernel_ms=kde_ms, peak_source=peaks["source"]) if kde_ms > 0 else None
    res = dict(metric="abm_agent_steps_per_s", value=value, unit="agent-steps/s", n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload="C2 ABMsolveensemble: %d realizations/GPU x NT=%d, n=%d, L=%d, M=%d, %d save points, "
                                    "KDE 89x89 sigma=0.1, Philox noise" % (R, NT, n, L, M, sp),
                           realizations_per_gpu=R, parallelism=f"ensemble-shard x{world}",
                           l2="flushed between timed passes (256 MiB write); working set is shared memory"),
               e2e=dict(value=e2e_value, unit="agent-steps/s", h2d_bytes_per_step=e["h2d_bytes_per_step"],
                        d2h_bytes_per_step=e["d2h_bytes_per_step"], ms_per_step=e["ms_per_step"], steps=e2e_steps,
                        host_bytes_written_per_step=e["host_bytes_written_per_step"], floor=floor,
                        api="smm_abm_ensemble (host buffers; us/zs/ys + ensemble sums returned as Float64; "
                            "us crosses PCIe as FP32 and is widened by host threads)" if e["host_widen"] else
                            "smm_abm_ensemble (host buffers; us/zs/ys + ensemble sums copied back as Float64)"),
               gpu_launches=launches, clocks=clocks, roofline=roofline, roofline_kde=roofline_kde,
               wall_ms_per_step=t["wall_s"] / args.steps * 1e3, ensemble_mass_check=e["mass_check"])
    return res, ctx, ens, flush


def abm_variants(args, ctx, ens, flush, rank, world):
    """Secondary ABM ensemble measurements: the end-to-end call with a pageable caller array / Float32 densities / means
    only, and configs[1] exactly as written (1 024 realizations over ALL GPUs of the job = 1 024 / world per GPU)."""
    import torch
    cfg = ens.cfg
    NT, n = cfg["NT"], cfg["n"]
    out = {}
    R = ens.R
    steps = max(1, min(args.steps, 3))
    for name, kw in (("e2e_pageable_caller_array", dict(mode="f64", pinned=False)),
                     ("e2e_float32_densities", dict(mode="f32", pinned=True)),
                     ("e2e_means_only", dict(mode="means"))):
        try:
            e = ens.time_e2e(steps, world, **kw)
            e["value"] = world * R * NT * n / (e["ms_per_step"] * 1e-3)
            e["unit"] = "agent-steps/s"
            out[name] = e
        except Exception as ex:
            out[name] = dict(error=f"{type(ex).__name__}: {ex}")
    # configs[1] as written: 1 024 realizations sharded over the job's GPUs (at 1 GPU also the 128-realization shard a
    # rank runs at 8 GPUs, to show the small-ensemble occupancy)
    shard_sizes = [C2["R"] // world] if world > 1 else [128]
    for Rs in shard_sizes:
        try:
            sub = AbmEnsemble(ctx, ens.stream, Rs, rank * Rs + 1, cfg)
            k = max(3, min(args.steps, 10))
            t = sub.time_device(k, 3, world, flush)
            e = sub.time_e2e(steps, world, mode="f64", pinned=True)
            em = sub.time_e2e(steps, world, mode="means")
            out[f"abm_C2_strong_{Rs}_per_gpu"] = dict(
                workload=f"configs[1] as written: {Rs * world} realizations over {world} GPU(s) = {Rs} per GPU",
                metric="abm_agent_steps_per_s", scaling="strong", value=world * Rs * NT * n * k / (t["dev_ms"] * 1e-3),
                ms_per_step=t["dev_ms"] / k, kernel_ms=t["kms"],
                e2e=dict(value=world * Rs * NT * n / (e["ms_per_step"] * 1e-3), ms_per_step=e["ms_per_step"],
                         d2h_bytes_per_step=e["d2h_bytes_per_step"]),
                e2e_means_only=dict(value=world * Rs * NT * n / (em["ms_per_step"] * 1e-3), ms_per_step=em["ms_per_step"],
                                    d2h_bytes_per_step=em["d2h_bytes_per_step"]))
            del sub
            torch.cuda.empty_cache()
        except Exception as ex:
            out[f"abm_C2_strong_{Rs}_per_gpu"] = dict(error=f"{type(ex).__name__}: {ex}")
    return out


def abm_one_call_multi_gpu(args, ens, world):
    """configs[1] through ONE call from ONE process (rank 0), every GPU of the job behind one context
    (smm_ctx_create_multi: one host thread per GPU, NCCL all-reduce of the sums inside the library) -- what the Julia
    shim's unchanged ABMsolveensemble(200, 1024) does with devices = all.  The other ranks idle at the next barrier."""
    import socialmediamodel_b200 as smm
    cfg = ens.cfg
    NT, n, R = cfg["NT"], cfg["n"], C2["R"]
    out = {}
    ctx = smm.Context(devices=list(range(world)))
    try:
        out["gpus"], out["nccl"] = ctx.gpu_count, ctx.nccl_active
        for name, kw in (("full_us_float64", dict(mode="f64", pinned=True)), ("float32_densities", dict(mode="f32", pinned=True)),
                         ("means_only", dict(mode="means"))):
            e = ens.time_e2e(3, 1, ctx=ctx, R=R, first=1, **kw)
            out[name] = dict(value=R * NT * n / (e["ms_per_step"] * 1e-3), unit="agent-steps/s", ms_per_step=e["ms_per_step"],
                             d2h_bytes_per_step=e["d2h_bytes_per_step"], mass_check=e["mass_check"])
    finally:
        ctx.close()
    out["workload"] = f"configs[1]: ABMsolveensemble(200, {R}) in ONE call, realizations sharded over {world} GPUs inside the library"
    return out


# ------------------------------------------------------------------------------------------
# C5: one large CSR population (HBM-bound kernel) -- secondary, 1 GPU
# ------------------------------------------------------------------------------------------
def run_abm_pop(args, ctx, n=10_000_000, L=64, k=16, steps=20, warmup=3):
    import torch

    import socialmediamodel_b200 as smm
    p, q = smm.ABMconstruct(), smm.parameters(n=n, L=L)
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-2.2, 2.2, (n, 2))
    state = np.where(rng.random(n) > 0.5, 1, -1).astype(np.int32)
    g = int(round(np.sqrt(L)))
    cx = np.minimum(g - 1, ((x0[:, 0] + 2.2) / 4.4 * g).astype(np.int64))
    cy = np.minimum(g - 1, ((x0[:, 1] + 2.2) / 4.4 * g).astype(np.int64))
    follow0 = (cy * g + cx).astype(np.int32)
    cnt = np.bincount(follow0, minlength=L)
    inf0 = np.stack([np.bincount(follow0, weights=x0[:, d], minlength=L) / cnt for d in range(2)], axis=1)
    media0 = np.array([[-1.0, -1.0], [1.0, 1.0]])
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    nb = torch.randint(0, n - 1, (n, k), device="cuda", generator=gen, dtype=torch.int32)
    me = torch.arange(n, device="cuda", dtype=torch.int32)[:, None]
    nb = nb + (nb >= me).to(torch.int32)                       # uniform over the other agents
    cols = torch.cat([nb, me], dim=1).contiguous()             # + the self-loop (Q4)
    colidx = cols.cpu().numpy().reshape(-1)
    rowptr = np.arange(0, n * (k + 1) + 1, k + 1, dtype=np.int32)
    del nb, cols
    pop = smm.ABMPopulation(q, p, x0, state, follow0, inf0, media0, (rowptr, colidx), seed=0, precision="f32", ctx=ctx)
    pop.step(warmup); ctx.sync()
    ms = []
    for _ in range(3):
        pop.step(steps); ctx.sync()
        ms.append(ctx.kernel_ms(0) / steps)
    ms_step = float(np.median(ms))
    st = pop.stats()
    got = pop.read(positions=False)
    pop.close()
    peaks = measured_peaks()
    bytes_per = 23 + 12 * (k + 1)                              # B = 23 + 12 deg
    ach = n * bytes_per / (ms_step * 1e-3) / 1e9
    # measured ceiling of the access pattern itself: (k+1) random 8-byte gathers per agent from the same 80 MB table
    # under the same persisting-L2 window, index rows streamed, nothing else (smm_microbench GATHER)
    try:
        gather_peak = ctx.microbench("gather", n)              # G gathers/s
    except Exception:
        gather_peak = None
    gathers = n * (k + 1) / (ms_step * 1e-3) / 1e9
    return dict(workload=f"C5 ABM population n={n}, L={L}, CSR out-degree {k}+self, Philox, FP32",
                metric="abm_agent_steps_per_s", value=n / (ms_step * 1e-3), ms_per_step=ms_step,
                fp64_reevaluation_rate=st["fp64_reevaluations"] / (st["steps"] * n),
                followers_conserved=bool(int(got["numbers"].sum()) == n),
                roofline=dict(kernel="pop_step_kernel<float>", bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s",
                              frac=ach / peaks["hbm_gbs"], traffic=NCU_TRAFFIC["pop_step"][0], traffic_source=NCU_TRAFFIC["pop_step"][1],
                              bytes_per_agent_step=bytes_per, peak_source=peaks["source"]),
                gather_roofline=dict(bound="l2-gather", achieved=gathers, peak=gather_peak, unit="Ggather/s",
                                     frac=gathers / gather_peak if gather_peak else None,
                                     note="peak = pure random float2 gather kernel, same table / window / degree, measured now"))


def run_abm_c1(ctx):
    """C1: ABMsolve() -- one realization, default paper parameters (n=250, L=4, M=2, NT=100), full history
    returned through the reference-shaped API; the oracle port of the same call timed beside it."""
    import socialmediamodel_b200 as smm
    from oracle import abm_oracle as A
    smm.ABMsolve(100, chosenseed=0, ctx=ctx)
    t0 = time.perf_counter()
    reps = 20
    for k in range(reps):
        xs, stateinfs, infs, meds, state, _ = smm.ABMsolve(100, chosenseed=k, ctx=ctx)
    gpu_ms = (time.perf_counter() - t0) / reps * 1e3
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-2.2, 2.2, (250, 2)); st = np.where(rng.random(250) > 0.5, 1, -1)
    f0, i0, m0 = A.abm_init(x0)
    t0 = time.perf_counter()
    A.abm_solve(100, x0, st, f0, i0, m0, rng.standard_normal((100, 250, 2)), rng.random((100, 250)), rng.random((100, 250)),
                A.default_p(), A.default_q())
    cpu_ms = (time.perf_counter() - t0) * 1e3
    return dict(workload="C1 ABMsolve(): 1 realization, n=250, L=4, M=2, NT=100, full history (host API, Float64 out)",
                metric="abm_agent_steps_per_s", value=250 * 100 / (gpu_ms * 1e-3), ms_per_call=gpu_ms,
                cpu_port_ms_per_call=cpu_ms, note="single CTA: latency-bound by design; ensembles are the throughput path")


# ------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle (port of the reference algorithm) on host cores
# ------------------------------------------------------------------------------------------
def _cpu_realization(k):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import abm_oracle as A
    from oracle import ensemble_oracle as E
    NT, n = C2["NT"], C2["n"]
    rng = np.random.default_rng(k)
    p, q = A.default_p(), A.default_q()
    x0 = rng.uniform(-2.2, 2.2, (n, 2)); state = np.where(rng.random(n) > 0.5, 1, -1)
    f0, i0, m0 = A.abm_init(x0)
    o = A.abm_solve(NT, x0, state, f0, i0, m0, rng.standard_normal((NT, n, 2)), rng.random((NT, n)),
                    rng.random((NT, n)), p, q)
    for t in E.savetimes_abm(NT, C2["savepoints"]):
        E.kde_snapshot(o["xs"][t], o["follows"][t], state, 4, p["X"], p["Y"], n)
    return NT * n


def _c_port_run(realizations, first=1):
    """R realizations of C2 with the plain-C restatement of the reference algorithm (oracle/abm_oracle.c:
    the reference's own loops incl. its non-separated KDE), OpenMP over realizations on all host cores."""
    from oracle import abm_oracle as A
    from oracle import c_oracle as CO
    from oracle import ensemble_oracle as E
    p, q = A.default_p(), A.default_q()
    st = E.savetimes_abm(C2["NT"], C2["savepoints"])
    t0 = time.perf_counter()
    tot, used = CO.ensemble_bench(C2["NT"], realizations, p, q, st, first=first, threads=host_threads())
    dt = time.perf_counter() - t0
    assert np.isfinite(tot)
    return realizations * C2["NT"] * C2["n"], dt, used


def julia_reference():
    """The reference itself (baseline/julia_bench.jl), when a `julia` binary and the reference checkout (SMM_REFERENCE_DIR)
    are present on the box -- they are not in the build image: returns None there."""
    import shutil
    ref = os.environ.get("SMM_REFERENCE_DIR")
    if not shutil.which("julia") or not ref or not os.path.isdir(ref):
        return None
    try:
        out = subprocess.run(["julia", "-t", "auto", f"--project={ref}", os.path.join(ROOT, "baseline", "julia_bench.jl"), ref, "16"],
                             capture_output=True, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        return dict(error=f"{type(e).__name__}: {e}")


def cpu_baseline_abm(realizations=None):
    cores = host_threads()
    try:
        _c_port_run(cores)                                    # warm-up (library load, thread pool)
        realizations = realizations or min(C2["R"], 64 * cores)   # ~10 s on 16 cores: the whole C2 ensemble
        done, dt, used = _c_port_run(realizations)
        return dict(value=done / dt, unit="agent-steps/s", cores=used, kind="port",
                    sample=f"{realizations} realizations of C2 (NT=200, n=250, KDE at 4 save points): C restatement of the "
                           f"reference's loops (oracle/abm_oracle.c, gcc -O3, Float64), OpenMP over realizations, {used} threads",
                    seconds=dt, julia_reference=julia_reference() or "julia absent (baseline/julia_bench.jl not run)")
    except (OSError, AssertionError, subprocess.CalledProcessError) as e:   # no compiler: NumPy restatement instead
        import multiprocessing as mp
        realizations = realizations or 12 * max(cores, 8)
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_realization, range(cores))
            t0 = time.perf_counter()
            done = sum(pool.map(_cpu_realization, range(realizations)))
            dt = time.perf_counter() - t0
        return dict(value=done / dt, unit="agent-steps/s", cores=cores, kind="port",
                    sample=f"{realizations} realizations of C2, NumPy Float64 oracle, one process per core ({cores}); "
                           f"C port unavailable: {type(e).__name__}", seconds=dt)


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which silently turned the
    reference arm into a single-thread run at N >= 2; the reference arm runs on rank 0 alone, so it takes every core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    # each step = the whole C2 ensemble (1 024 realizations), i.e. the native arm's per-GPU step: same config.  On 16
    # cores that is ~9.5 s per step; --steps is capped so that the arm ends within a few minutes.
    per_step = C2["R"]
    budget_s = float(os.environ.get("SMM_REF_BUDGET_S", "150"))
    from oracle import abm_oracle as A
    from oracle import c_oracle as CO
    from oracle import ensemble_oracle as E
    p, q = A.default_p(), A.default_q()
    st = E.savetimes_abm(C2["NT"], C2["savepoints"])

    def run(R, first):
        t0 = time.perf_counter()
        tot, used = CO.ensemble_bench(C2["NT"], R, p, q, st, first=first, threads=threads)
        assert np.isfinite(tot)
        return time.perf_counter() - t0, used

    t_w, used = run(4 * threads, 1)                      # warm-up (library load, thread pool) + a rate estimate
    est_step = t_w * per_step / (4 * threads)
    steps = max(1, min(args.steps, int(budget_s / max(est_step, 1e-9))))
    dt = 0.0
    for s in range(steps):
        t, used = run(per_step, 1 + s * per_step)
        dt += t
    v = steps * per_step * C2["NT"] * C2["n"] / dt
    sample = (f"each step = {per_step} realizations of C2 (NT=200, n=250, KDE at 4 save points) on {used} host threads: C "
              "restatement of the reference's loops (oracle/abm_oracle.c, Float64; Julia is not installed, the reference "
              f"itself cannot run here); {steps} of the requested {args.steps} steps timed (time cap {budget_s:.0f} s)")
    print(json.dumps(dict(impl="reference", metric="abm_agent_steps_per_s", value=v, unit="agent-steps/s",
                          n_gpus=args.gpus, steps=steps, warmup=args.warmup, ms_per_step=dt / steps * 1e3,
                          higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                          config=dict(workload="C2 ABMsolveensemble: %d realizations x NT=%d, n=%d, L=%d, M=%d, %d save points, "
                                               "KDE 89x89 sigma=0.1" % (per_step, C2["NT"], C2["n"], C2["L"], C2["M"], C2["savepoints"]),
                                      realizations_per_gpu=per_step, sample=sample),
                          cpu_baseline=dict(value=v, unit="agent-steps/s", cores=used, kind="port", sample=sample),
                          e2e=dict(value=v, unit="agent-steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                          gpu_launches=0)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--realizations", type=int, default=0, help="realizations per GPU (default 1024)")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads and the CPU baseline")
    ap.add_argument("--cpu-baseline-only", action="store_true", help="(internal) print the cpu_baseline object and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.cpu_baseline_only:
        print(json.dumps(cpu_baseline_abm()))
        return
    args.warmup = max(args.warmup, 3)
    rank, world, local = dist_setup(args.gpus)
    res, ctx, ens, flush = run_abm_ensemble(args, rank, world, local)
    if not args.no_extra:
        extra = {}

        def attempt(name, fn, everyone=True):
            """a secondary workload must not lose the headline line; `everyone`: all ranks take part (collectives inside)"""
            if not everyone and rank != 0:
                return
            try:
                r = fn()
                if isinstance(r, dict) and name is None:
                    extra.update(r)
                else:
                    extra[name] = r
            except Exception as e:
                extra[name or fn.__name__] = dict(error=f"{type(e).__name__}: {e}")

        attempt(None, lambda: abm_variants(args, ctx, ens, flush, rank, world))
        from bench_pde import run_pde_c4_job, run_pde_workloads
        attempt("pde_C4", lambda: run_pde_c4_job(ctx, rank, world))          # every rank: its shard of the sweep + all-reduce
        if world == 1:
            attempt("abm_C1", lambda: run_abm_c1(ctx))
            attempt("abm_population_C5", lambda: run_abm_pop(args, ctx))
            attempt(None, lambda: run_pde_workloads(ctx))
        else:
            # rank 0 drives every GPU of the job from one process; the other ranks must leave their GPUs idle meanwhile.
            # They wait on a gloo (CPU) barrier: an NCCL barrier would park a spinning kernel on their GPU and the two
            # processes' contexts would time-slice it with the measured work
            import torch
            import torch.distributed as dist
            cpu_group = dist.new_group(backend="gloo")
            barrier(world)
            torch.cuda.synchronize()
            attempt("abm_C2_one_call", lambda: abm_one_call_multi_gpu(args, ens, world), everyone=False)
            dist.barrier(group=cpu_group)
        res["workloads"] = extra
        # The CPU baseline runs last and in a child process: ten seconds of all-core work, and the oracle, stay out of the
        # process and of the time in which the GPU path and the end-to-end call are measured.
        cpu = None
        if world == 1:
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only"], capture_output=True, text=True,
                                     timeout=900)
                cpu = json.loads(out.stdout.strip().splitlines()[-1])
            except Exception as e:
                cpu = dict(error=f"{type(e).__name__}: {e}")
        res["cpu_baseline"] = cpu
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
