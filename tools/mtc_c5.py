"""Per-GPU slice of config 5 (n = 2M here, 10 + 10 terms, 128 chains), float32: tensor-core
two-pass form against the CUDA-core sweep (PTM_MAIN_TC=0)."""
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
res = {}
for env in ("", "PTM_MAIN_TC=0"):
    if env: os.environ["PTM_MAIN_TC"] = "0"
    wl = bw.build("c5_n50M_10loc_10scale_128chains", dtype=np.float32, n_override=2_000_000)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).cuda()
    lpv = torch.empty(p.shape[0], dtype=p.dtype, device="cuda"); g = torch.empty_like(p)
    a = t_ms(lambda: m.log_prob_and_grad(p, lpv, g)); b = t_ms(lambda: m.log_prob(p, lpv))
    res[env] = (lpv.clone(), g.clone())
    print(f"[{env or 'default'}] {m.last_main_route()} logp+grad {a:.2f} ms  logp {b:.2f} ms  ({2e6 * 128 / a * 1e3:.3e} evals/s)", flush=True)
    os.environ.pop("PTM_MAIN_TC", None)
    del m, wl
lp1, g1 = res[""]; lp0, g0 = res["PTM_MAIN_TC=0"]
print("logp rel", float(((lp1 - lp0).abs() / lp0.abs()).max()), "grad rel inf", float(((g1 - g0).abs().amax(1) / g0.abs().amax(1)).max()))
