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
included.  Secondary workloads (C5 population kernel, C3/C4 PDE) are reported under "workloads".
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
def run_abm_ensemble(args, rank, world, local):
    import ctypes as C

    import torch

    import socialmediamodel_b200 as smm
    from socialmediamodel_b200 import _lib

    cfg = dict(C2)
    if args.realizations:
        cfg["R"] = args.realizations
    R, NT, n, L, M, sp = cfg["R"], cfg["NT"], cfg["n"], cfg["L"], cfg["M"], cfg["savepoints"]
    p, q = smm.ABMconstruct(), smm.parameters(n=n, L=L)
    ctx = smm.Context(local)
    stream = torch.cuda.Stream()             # a real stream (the legacy default stream cannot be graph-captured)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    Nx, Ny = p.X.shape
    savesteps = smm.savetimes_abm(NT, sp)
    per_real = Nx * Ny * 2 * L * sp
    us_dev = torch.empty(R * per_real, dtype=torch.float32, device="cuda")
    sum_us = torch.zeros(per_real, dtype=torch.float64, device="cuda")
    sum_zs = torch.zeros(2 * 2 * sp, dtype=torch.float64, device="cuda")
    sum_ys = torch.zeros(2 * L * sp, dtype=torch.float64, device="cuda")
    numbers = torch.zeros(M * L * (NT + 1), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    spar, grid = _lib.make_params(q, p), _lib.make_grid(p)
    first = rank * R + 1                                               # seeds k = 1..N (ensemble.jl:25)

    def one_pass():
        rc = ctx.lib.smm_abm_ensemble_dev(ctx.h, C.byref(spar), 0, NT, R, first, 0, None, None, sp,
                                          _lib.ptr(savesteps), cfg["kde_sigma"], C.byref(grid),
                                          C.c_void_p(us_dev.data_ptr()), C.c_void_p(sum_us.data_ptr()),
                                          C.c_void_p(sum_zs.data_ptr()), C.c_void_p(sum_ys.data_ptr()),
                                          C.c_void_p(numbers.data_ptr()))
        ctx.check(rc, "smm_abm_ensemble_dev")
        if world > 1:   # ensemble mean over all GPUs (plotting.jl:361-363): the path's only exchange
            import torch.distributed as dist
            dist.all_reduce(sum_us); dist.all_reduce(sum_zs); dist.all_reduce(sum_ys)

    sampler = ClockSampler(local); sampler.start()   # started before the warm-up so it is sampling by the timed region
    for _ in range(args.warmup):
        one_pass(); flush.fill_(1)
    barrier(world)
    l0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kms = {"abm": [], "kde": [], "reduce": []}
    barrier(world)
    sampler.mark()
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        ev[k][0].record(stream)
        one_pass()
        ev[k][1].record(stream)
        flush.fill_(k & 1)                       # L2 flush between timed passes (outside the event pair)
        ctx.sync()                               # collects the per-kernel CUDA-event times of this pass
        kms["abm"].append(ctx.kernel_ms(0)); kms["kde"].append(ctx.kernel_ms(1)); kms["reduce"].append(ctx.kernel_ms(2))
    barrier(world)
    t_wall = time.perf_counter() - t_wall0
    launches = ctx.launches - l0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = max_over_ranks(dev_ms, world)
    agent_steps = world * R * NT * n * args.steps
    value = agent_steps / (dev_ms * 1e-3)

    # ---- e2e: host-buffer C-ABI call (what the Julia shim's ABMsolveensemble makes) ------------
    out_us = ctx.pinned_empty((Nx, Ny, 2, L, sp, R))
    zs = np.empty((2, 2, sp, R), order="F"); ys = np.empty((2, L, sp, R), order="F")
    s_us = np.empty((Nx, Ny, 2, L, sp), order="F"); s_zs = np.empty((2, 2, sp), order="F"); s_ys = np.empty((2, L, sp), order="F")
    ns = _lib.smm_noise(); ns.mode = 0; ns.seed = 0

    def e2e_pass():
        rc = ctx.lib.smm_abm_ensemble(ctx.h, C.byref(spar), 0, NT, R, first, None, None, None, None, None, None, None,
                                      C.byref(ns), sp, _lib.ptr(savesteps), cfg["kde_sigma"], C.byref(grid),
                                      _lib.ptr(out_us), _lib.ptr(zs), _lib.ptr(ys), _lib.ptr(s_us), _lib.ptr(s_zs),
                                      _lib.ptr(s_ys), None)
        ctx.check(rc, "smm_abm_ensemble")

    e2e_steps = max(1, min(args.steps, 5))
    e2e_pass()
    barrier(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass()
    barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    e2e_value = world * R * NT * n * e2e_steps / e2e_s
    # the per-realization densities cross PCIe as FP32 and are widened into the Float64 host array by the
    # library's host threads (csrc/host_widen.cpp); everything else is copied as Float64
    hw = os.environ.get("SMM_HOST_WIDEN")
    widen = hw == "1" or (hw != "0" and (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", "1")) >= 12)
    d2h = (out_us.nbytes // 2 if widen else out_us.nbytes) + zs.nbytes + ys.nbytes + s_us.nbytes + s_zs.nbytes + s_ys.nbytes
    h2d = C.sizeof(_lib.smm_params) + C.sizeof(_lib.smm_grid) + savesteps.nbytes
    mean_check = float(s_us.sum() * p.dV / (R * sp))
    ctx.free_pinned(out_us)

    peaks = measured_peaks()
    abm_ms = float(np.mean(kms["abm"]))
    pairs = R * NT * n * n                                              # ordered pairs per launch
    sym = os.environ.get("SMM_ABM_SYM", "1") != "0"
    mufu_per_pair = 1 if sym else 2    # sqrt + ex2 per evaluated pair; the symmetric kernel evaluates unordered pairs
    mufu_peak = 148 * 16 * peaks["sm_max_mhz"] * 1e6 / mufu_per_pair / 1e9   # G ordered pairs/s, 16 MUFU/clk/SM
    # issue-slot view (ncu, profiles/r1_abm_dense_sym_*): warp instructions per realization-step, 4 issue slots/clk/SM
    winst = 19860 if sym else 27700
    issue_ms = R * NT * winst / (148 * 4 * peaks["sm_max_mhz"] * 1e6) * 1e3
    roofline = dict(kernel="abm_dense_kernel<float,sym>" if sym else "abm_dense_kernel<float>", bound="mufu",
                    achieved=pairs / (abm_ms * 1e-3) / 1e9, peak=mufu_peak, unit="Gpair/s",
                    frac=pairs / (abm_ms * 1e-3) / 1e9 / mufu_peak, traffic=None,
                    note="state lives in shared memory; not HBM-bound (SURVEY 8d). Algorithmic work: 2 MUFU (sqrt, ex2) per "
                         "UNORDERED pair = 1 per ordered pair (w_ij = w_ji), "
                         f"peak = 148 SM x 16 MUFU/clk x {peaks['sm_max_mhz']:.0f} MHz (derived, nominal clock); "
                         "the kernel is co-limited by instruction issue (issue_frac)",
                    issue_frac=issue_ms / abm_ms, warp_inst_per_realization_step=winst,
                    kernel_ms=abm_ms, share_of_step=abm_ms * args.steps / dev_ms if world == 1 else None)
    kde_ms = float(np.mean(kms["kde"]))
    kde_bytes = R * per_real * 4
    roofline_kde = dict(kernel="kde_kernel<float>", bound="hbm", achieved=kde_bytes / (kde_ms * 1e-3) / 1e9,
                        peak=peaks["hbm_gbs"], unit="GB/s", frac=kde_bytes / (kde_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                        traffic=None, kernel_ms=kde_ms, peak_source=peaks["source"])
    res = dict(metric="abm_agent_steps_per_s", value=value, unit="agent-steps/s", n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling="weak",
               vs_baseline=None, dtype="f32", data="synthetic",
               config=dict(workload="C2 ABMsolveensemble: %d realizations/GPU x NT=%d, n=%d, L=%d, M=%d, %d save points, "
                                    "KDE 89x89 sigma=0.1, Philox noise" % (R, NT, n, L, M, sp),
                           realizations_per_gpu=R, parallelism=f"ensemble-shard x{world}",
                           l2="flushed between timed passes (256 MiB write); working set is shared memory"),
               e2e=dict(value=e2e_value, unit="agent-steps/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                        ms_per_step=e2e_s / e2e_steps * 1e3, steps=e2e_steps,
                        host_bytes_written_per_step=out_us.nbytes + zs.nbytes + ys.nbytes + s_us.nbytes + s_zs.nbytes + s_ys.nbytes,
                        api="smm_abm_ensemble (host buffers; us/zs/ys + ensemble sums returned as Float64; "
                            "us crosses PCIe as FP32 and is widened by host threads)" if widen else
                            "smm_abm_ensemble (host buffers; us/zs/ys + ensemble sums copied back as Float64)"),
               gpu_launches=launches, clocks=clocks, roofline=roofline, roofline_kde=roofline_kde,
               wall_ms_per_step=t_wall / args.steps * 1e3, ensemble_mass_check=mean_check)
    return res, ctx


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
    return dict(workload=f"C5 ABM population n={n}, L={L}, CSR out-degree {k}+self, Philox, FP32",
                metric="abm_agent_steps_per_s", value=n / (ms_step * 1e-3), ms_per_step=ms_step,
                fp64_reevaluation_rate=st["fp64_reevaluations"] / (st["steps"] * n),
                followers_conserved=bool(int(got["numbers"].sum()) == n),
                roofline=dict(kernel="pop_step_kernel<float>", bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s",
                              frac=ach / peaks["hbm_gbs"], traffic=None, bytes_per_agent_step=bytes_per,
                              peak_source=peaks["source"]))


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
    tot, used = CO.ensemble_bench(C2["NT"], realizations, p, q, st, first=first)
    dt = time.perf_counter() - t0
    assert np.isfinite(tot)
    return realizations * C2["NT"] * C2["n"], dt, used


def cpu_baseline_abm(realizations=None):
    cores = os.cpu_count() or 1
    try:
        _c_port_run(cores)                                    # warm-up (library load, thread pool)
        realizations = realizations or min(C2["R"], 64 * cores)   # ~10 s on 16 cores: the whole C2 ensemble
        done, dt, used = _c_port_run(realizations)
        return dict(value=done / dt, unit="agent-steps/s", cores=used, kind="port",
                    sample=f"{realizations} realizations of C2 (NT=200, n=250, KDE at 4 save points): C restatement of the "
                           f"reference's loops (oracle/abm_oracle.c, gcc -O3, Float64), OpenMP over realizations, {used} threads",
                    seconds=dt)
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = 4 * max(cores, 8)
    for _ in range(min(args.warmup, 1)):
        _c_port_run(cores)
    done, dt, used = 0, 0.0, cores
    for s in range(args.steps):
        d, t, used = _c_port_run(per_step, first=1 + s * per_step)
        done += d; dt += t
    v = done / dt
    sample = (f"each step = {per_step} realizations of C2 (NT=200, n=250, KDE at 4 save points) on {used} host threads: C "
              "restatement of the reference's loops (oracle/abm_oracle.c, Float64; Julia is not installed, the reference "
              "itself cannot run here)")
    print(json.dumps(dict(impl="reference", metric="abm_agent_steps_per_s", value=v, unit="agent-steps/s",
                          n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=dt / args.steps * 1e3,
                          higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                          config=dict(workload="C2 ABMsolveensemble (bounded sample)", sample=sample),
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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    args.warmup = max(args.warmup, 3)
    cpu = None
    if int(os.environ.get("WORLD_SIZE", "1")) == 1 and not args.no_extra:
        cpu = cpu_baseline_abm()          # before CUDA is initialised in this process (fork-safe)
    rank, world, local = dist_setup(args.gpus)
    res, ctx = run_abm_ensemble(args, rank, world, local)
    if rank == 0 and world == 1 and not args.no_extra:
        extra = {}
        for name, fn in (("abm_C1", lambda: run_abm_c1(ctx)), ("abm_population_C5", lambda: run_abm_pop(args, ctx))):
            try:
                extra[name] = fn()
            except Exception as e:  # a secondary workload must not lose the headline line
                extra[name] = dict(error=f"{type(e).__name__}: {e}")
        try:
            from bench_pde import run_pde_workloads
            extra.update(run_pde_workloads(ctx))
        except ImportError:
            pass
        res["workloads"] = extra
        res["cpu_baseline"] = cpu
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
