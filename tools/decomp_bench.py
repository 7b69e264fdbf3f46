"""How the c3 sweep time splits between the additive predictors and the point-wise part:
n = 1M, 1024 chains, T = 0 / 2 / 6 / 10 smooth terms (half loc, half scale), logp+grad and
logp only, both precisions.  CUDA events, 3 warm-ups, mean of 5."""
import json, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
import liesel_ptm_b200 as lp

def t_ms(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {}
n, C = 1_000_000, 1024
for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
    for T in (0, 2, 6, 10):
        y, X = bw.make_data(n, T // 2, T - T // 2, seed=0, dtype=dt)
        m = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=(dt == np.float32), device="cuda")
        for t in range(T):
            tm = lp.term.f_ig(lp.ps(X[t], nbases=20, xname=f"x{t}", dtype=dt))
            if t < T // 2: m.loc += tm
            else: m.scale += tm
        p = torch.from_numpy(bw.make_states(m, C, seed=1).astype(dt)).cuda()
        m.build()
        lpv = torch.empty(C, dtype=p.dtype, device="cuda"); g = torch.empty_like(p)
        a = t_ms(lambda: m.log_prob_and_grad(p, lpv, g)); b = t_ms(lambda: m.log_prob(p, lpv))
        out[f"{tag}_T{T}"] = {"logp_grad_ms": a, "logp_ms": b}
        print(tag, T, f"grad {a:.2f} ms   logp {b:.2f} ms", flush=True)
        del m, p, g
        torch.cuda.empty_cache()
json.dump(out, open("gpurun_out/r2_decomp.json", "w"), indent=1)
