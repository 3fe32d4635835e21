"""Time the PDE step with the generic FFT kernels and the register-resident ones (both CTA shapes):
   python tools/pde_fft_compare.py            -> one JSON line per variant and workload"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import bench_pde
    import socialmediamodel_b200 as smm
    ctx = smm.Context.default()
    out = dict(variant=os.environ.get("VARIANT"))
    out["c3_us"] = bench_pde.run_pde_c3(ctx, steps=200)["us_per_step"]
    out["s177x16_ms"] = bench_pde.run_pde_sweep(ctx, dx=0.025, B=16, steps=4)["ms_per_step"]
    out["s353x16_ms"] = bench_pde.run_pde_sweep(ctx, dx=0.0125, B=16, steps=3)["ms_per_step"]
    out["s353x64_ms"] = bench_pde.run_pde_sweep(ctx, dx=0.0125, B=64, steps=3)["ms_per_step"]
    print(json.dumps(out))
else:
    for name, env in (("generic", {"SMM_FFT_FAST": "0"}), ("fast4", {"SMM_FFT_PAIRS": "4"}), ("fast8", {"SMM_FFT_PAIRS": "8"}),
                      ("fast4_chunk8", {"SMM_FFT_PAIRS": "4", "SMM_FFT_CHUNK_MB": "36"}),
                      ("fast4_nochunk", {"SMM_FFT_PAIRS": "4", "SMM_FFT_CHUNK_MB": "4000"})):
        e = dict(os.environ, VARIANT=name, **env)
        r = subprocess.run([sys.executable, __file__, "child"], env=e, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + name + ": " + r.stderr[-800:], flush=True)
