"""PDE workloads of bench.py (secondary lines under "workloads"): C3 = PDEsolve on the default
89x89x2x4 grid (one instance; launch/latency-bound, reported as us per step), and a bounded
sample of C4 = a parameter sweep of PDE instances on refined grids (HBM-bound stage kernel)."""
import numpy as np


def _peaks():
    from bench import measured_peaks
    return measured_peaks()


def _sweep_qs(smm, B):
    """deterministic parameter lattice of C4."""
    qs = []
    for a in (0.5, 1.0, 2.0, 3.0):
        for b in (1.0, 2.0, 4.0, 6.0):
            for c in (1.0, 2.0, 3.0, 4.0):
                for eta in (1.0, 15.0):
                    for sigma in (0.25, 0.5, 0.75, 1.0):
                        qs.append(smm.parameters(a=a, b=b, c=c, eta=eta, sigma=sigma))
    return qs[:B]


def run_pde_c3(ctx, steps=400, dt=1e-3):
    import socialmediamodel_b200 as smm
    p, q = smm.PDEconstruct(), smm.parameters()
    uzy0, q2 = smm.constructinitial("4inf", (p, q))
    pb = smm.PDEBatch(p, [q2], *uzy0.x, precision="f32", ctx=ctx)
    pb.step(dt, 20); ctx.sync()
    ms = []
    l0 = ctx.launches
    for _ in range(3):
        pb.step(dt, steps); ctx.sync()
        ms.append(ctx.kernel_ms(4) / steps)
    launches_per_step = (ctx.launches - l0) / (3 * steps)
    us_step = float(np.median(ms)) * 1e3
    u, z, y = pb.read()
    pb.close()
    cells = p.N_x * p.N_y * 2 * 4
    cluster = launches_per_step < 1
    how = ("all steps of a call in ONE launch: the instance lives in the shared memory of a 16-CTA thread-block cluster (pde_cluster.cu)"
           if cluster else "16 kernels per step in a CUDA graph")
    return dict(workload=f"C3 PDEsolve 89x89x2x4, uniforminit, RK4 dt=1e-3, one instance, {how}",
                metric="pde_cell_steps_per_s", value=cells / (us_step * 1e-6), us_per_step=us_step,
                launches_per_step=launches_per_step, mass=float(u.sum() * p.dV), finite=bool(np.isfinite(u).all()),
                note="63 368 cells: latency-bound, not HBM-bound (SURVEY 8d); SMM_PDE_CLUSTER=0 gives the kernel chain (90 us per step)")


def run_pde_sweep(ctx, dx=0.025, B=64, steps=10, dt=None):
    import socialmediamodel_b200 as smm
    p = smm.PDEconstruct(dx=dx)
    qs = _sweep_qs(smm, B)
    uzy0, _ = smm.constructinitial("4inf", (p, qs[0]))
    dt = dt or min(smm.default_dt(p, q) for q in qs)
    pb = smm.PDEBatch(p, qs, *uzy0.x, precision="f32", shared_init=True, ctx=ctx)
    pb.step(dt, 2); ctx.sync()
    ms = []
    for _ in range(3):
        pb.step(dt, steps); ctx.sync()
        ms.append(ctx.kernel_ms(4) / steps)
    ms_step = float(np.median(ms))
    u, z, y = pb.read(want_u=False)
    pb.close()
    cells = p.N_x * p.N_y * 2 * 4 * B
    peaks = _peaks()
    ach = cells * 76 / (ms_step * 1e-3) / 1e9          # 19 B per cell per RK stage
    return dict(workload=f"C4 sample: {B} PDE instances {p.N_x}x{p.N_y}x2x4, parameter lattice, RK4 dt={dt:.2e}",
                metric="pde_cell_steps_per_s", value=cells / (ms_step * 1e-3), ms_per_step=ms_step,
                finite=bool(np.isfinite(y).all()),
                roofline=dict(kernel="pde_stage_kernel<float,4> + fft_{rows_fwd,cols,rows_inv}<float>", bound="hbm", achieved=ach,
                              peak=peaks["hbm_gbs"], unit="GB/s", frac=ach / peaks["hbm_gbs"], traffic=None,
                              bytes_per_cell_step=76, peak_source=peaks["source"],
                              note="whole RK4 step: 4 x (3 FFT kernels for the interaction force + fused stage kernel); "
                                   "76 B/cell-step counts the stencil/RK traffic only (SURVEY 8d)"))


def run_pde_c4_job(ctx, rank, world, B=64, steps=3, dx=0.0125):
    """C4 (configs[3]): the 512-instance parameter sweep at the 4x refined grid, 64 instances per GPU.  Rank r integrates
    lattice points r*64 .. r*64+63 as one device-resident batch (8 GPUs = the whole lattice; fewer GPUs = its first
    world*64 points: weak scaling), every rank times its own RK4 steps with CUDA events, the job's time is the max over
    ranks, and the ensemble-mean state (plotting.jl:361-363) is formed with one NCCL all-reduce of the per-rank sums --
    the path's only exchange (ensemble.jl:58-85: the instance loop :69-70 is embarrassingly parallel)."""
    import torch

    import socialmediamodel_b200 as smm
    from bench import barrier, max_over_ranks
    p = smm.PDEconstruct(dx=dx)
    qs_all = _sweep_qs(smm, 512)
    qs = [qs_all[(rank * B + k) % 512] for k in range(B)]
    uzy0, _ = smm.constructinitial("4inf", (p, qs[0]))
    dt = min(smm.default_dt(p, q) for q in qs_all)
    pb = smm.PDEBatch(p, qs, *uzy0.x, precision="f32", shared_init=True, ctx=ctx)
    pb.step(dt, 2); ctx.sync()
    barrier(world)
    ms = []
    for _ in range(3):
        pb.step(dt, steps); ctx.sync()
        ms.append(ctx.kernel_ms(4) / steps)
        barrier(world)
    ms_step = max_over_ranks(float(np.median(ms)), world)
    # ensemble mean over the whole job: per-rank sums on the device, one all-reduce
    per = p.N_x * p.N_y * 2 * 4
    su = torch.zeros(per, dtype=torch.float64, device="cuda"); sz = torch.zeros(4, dtype=torch.float64, device="cuda")
    sy = torch.zeros(8, dtype=torch.float64, device="cuda")
    import ctypes as C
    ctx.check(ctx.lib.smm_pde_batch_sum_dev(ctx.h, pb.h, C.c_void_p(su.data_ptr()), C.c_void_p(sz.data_ptr()),
                                            C.c_void_p(sy.data_ptr())), "smm_pde_batch_sum_dev")
    ctx.sync()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(su); dist.all_reduce(sz); dist.all_reduce(sy)
    t1.record(); torch.cuda.synchronize()
    allreduce_ms = t0.elapsed_time(t1)
    mean_mass = float(su.sum().item()) * p.dV / (B * world)
    pb.close()
    cells = p.N_x * p.N_y * 2 * 4 * B * world
    peaks = _peaks()
    ach = cells / world * 76 / (ms_step * 1e-3) / 1e9      # per GPU: 19 B per cell per RK stage
    from bench import NCU_TRAFFIC
    stage_traffic = NCU_TRAFFIC["pde_stage2_per16"][0] * B / 16 if p.N_x == 353 else None
    return dict(workload=f"C4 parameter sweep: {B} PDE instances per GPU x {world} GPU(s), {p.N_x}x{p.N_y}x2x4, RK4 dt={dt:.2e}, "
                         "lattice points sharded by rank, ensemble mean all-reduced",
                metric="pde_cell_steps_per_s", scaling="weak", n_gpus=world, value=cells / (ms_step * 1e-3), ms_per_step=ms_step,
                instances_per_gpu=B, mean_all_reduce_ms=allreduce_ms, ensemble_mean_mass=mean_mass,
                roofline=dict(kernel="pde_stage_kernel<float,4> + FFT convolution kernels", bound="hbm", achieved=ach,
                              peak=peaks["hbm_gbs"], unit="GB/s (per GPU)", frac=ach / peaks["hbm_gbs"],
                              traffic=stage_traffic, traffic_note="DRAM bytes of ONE stage-kernel launch (4 per step), ncu: "
                              + NCU_TRAFFIC["pde_stage2_per16"][1],
                              bytes_per_cell_step=76, peak_source=peaks["source"],
                              stage_kernel=dict(share_of_step=0.446, frac_of_hbm=0.55, note="ncu launch list at 64 instances "
                                                "(profiles/r2_launches_pde_353x64.csv): stage kernel 340 us of 763 us per RK stage = "
                                                "1.21 GB / 340 us = 3.57 TB/s; FFT kernels 414 us (shared-memory-bound, not HBM)"),
                              note="whole RK4 step per GPU: 4 x (FFT evaluation of the interaction force + fused stage kernel); "
                                   "76 B/cell-step counts the stencil/RK traffic only (SURVEY 8d)"))


def run_pde_workloads(ctx):
    out = {}
    for name, fn in (("pde_C3", lambda: run_pde_c3(ctx)),
                     ("pde_sweep_177", lambda: run_pde_sweep(ctx, dx=0.025, B=16, steps=4))):
        try:
            out[name] = fn()
        except Exception as e:
            out[name] = dict(error=f"{type(e).__name__}: {e}")
    return out


if __name__ == "__main__":
    import json

    import socialmediamodel_b200 as smm
    print(json.dumps(run_pde_workloads(smm.Context.default()), indent=1))
