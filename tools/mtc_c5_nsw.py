"""Stager-warp count (PTM_MTC_NSW) on a per-GPU slice of config 5 (n = 2M, 10 + 10 terms, 128 chains,
float32, several-pass tensor-core form) and on the c4 term structure (10 x 30 + 30, 256 chains)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
def t_ms(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for nsw in sys.argv[1:] or ["0"]:
    os.environ["PTM_MTC_NSW"] = nsw
    for name, kw in (("c5_n50M_10loc_10scale_128chains", dict(n_override=2_000_000)), ("c4_n1M_10loc_1scale_nb30_256chains", {})):
        wl = bw.build(name, dtype=np.float32, **kw)
        m = wl.model.build()
        p = torch.from_numpy(wl.params).cuda()
        lpv = torch.empty(p.shape[0], dtype=p.dtype, device="cuda"); g = torch.empty_like(p)
        a = t_ms(lambda: m.log_prob_and_grad(p, lpv, g)); b = t_ms(lambda: m.log_prob(p, lpv))
        print(f"[NSW={nsw}] {name[:2]} {m.last_main_route()} logp+grad {a:.2f} ms  logp {b:.2f} ms", flush=True)
        del m, wl, p, g
        torch.cuda.empty_cache()
