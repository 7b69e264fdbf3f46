"""c3 variant of SURVEY.md 8(d): the five scale terms use the SAME covariates as the five
location terms (11 distinct columns -> 6).  ptm_problem_create detects the pairing and the
sweep stages / gathers every design once (PTM_NO_SHARED=1 turns that off); the algorithmic
bytes per evaluation drop to w * (1 + 5) = 48 B (f64)."""
import json, sys
sys.path.insert(0, '.')
import numpy as np, torch
import liesel_ptm_b200 as lp
import bench_workload as bw

n, C = 1_000_000, 1024
y, X = bw.make_data(n, 5, 5, seed=0)
model = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=False)
for t in range(10):
    tm = lp.term.f_ig(lp.ps(X[t % 5], nbases=20, xname=f"x{t % 5}"), fname="f" if t < 5 else "g")
    if t < 5: model.loc += tm
    else: model.scale += tm
model.build()
p = torch.from_numpy(bw.make_states(model, C, seed=1)).cuda()
lpv = torch.empty(C, dtype=torch.float64, device="cuda"); g = torch.empty_like(p)
for _ in range(3): model.log_prob_and_grad(p, lpv, g)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): model.log_prob_and_grad(p, lpv, g)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
rec = {"variant": "c3 with loc and scale sharing five covariates", "ms": ms, "evals_per_s": n * C / ms * 1e3,
       "bytes_alg_per_eval": 48, "frac_of_hbm_roofline": n * C / ms * 1e3 * 48 / 6560.3e9,
       "finite": bool(torch.isfinite(lpv).all())}
print(json.dumps(rec))
json.dump(rec, open("gpurun_out/r1_c3_shared_covariates.json", "w"), indent=1)
