"""A/B of create-time knobs on the c3 float sweep: python tools/mtc_ab.py "K=V,K=V" ... ; prints ms."""
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
dt = np.float32
for env in sys.argv[1:] or [""]:
    for kv in env.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", dtype=dt)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).cuda()
    lpv = torch.empty(p.shape[0], dtype=p.dtype, device="cuda"); g = torch.empty_like(p)
    a = t_ms(lambda: m.log_prob_and_grad(p, lpv, g)); b = t_ms(lambda: m.log_prob(p, lpv))
    print(f"[{env or 'default'}] {m.last_main_route()} logp+grad {a:.2f} ms  logp {b:.2f} ms", flush=True)
    for kv in env.split(","):
        if "=" in kv: os.environ.pop(kv.split("=")[0])
    del m, wl, p, g
    torch.cuda.empty_cache()
