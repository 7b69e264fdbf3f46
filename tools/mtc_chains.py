"""c3 float32 step time against the number of chains on one GPU (the per-GPU share of the
strong-scaling runs): python tools/mtc_chains.py [ENV=V,...]"""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
for kv in (sys.argv[1] if len(sys.argv) > 1 else "").split(","):
    if "=" in kv:
        k, v = kv.split("="); os.environ[k] = v
def t_ms(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for C in (128, 256, 512, 1024):
    wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=C, dtype=np.float32)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).cuda()
    lpv = torch.empty(C, dtype=p.dtype, device="cuda"); g = torch.empty_like(p)
    a = t_ms(lambda: m.log_prob_and_grad(p, lpv, g))
    print(f"C={C}: {a:.2f} ms  -> {1e6 * C / a * 1e3:.3e} evals/s  ({a / C * 1024:.2f} ms per 1024 chains)", flush=True)
    del m, wl, p, g; torch.cuda.empty_cache()
