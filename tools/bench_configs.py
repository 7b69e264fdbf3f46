"""Secondary measurements (not the bench.py contract line): the other BASELINE configs and
the chain-batched IWLS information.  CUDA events, 3 warm-ups, mean of 5."""
import sys, json
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw

def timeit(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

out = {}
for name in ["c1_readme_n100", "c2_n100k_3loc_1scale_64chains", "c3_n1M_5loc_5scale_nb20_1024chains",
             "c4_n1M_10loc_1scale_nb30_256chains"]:
    wl = bw.build(name, dtype=np.float64)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).cuda()
    C = p.shape[0]
    lp = torch.empty(C, dtype=torch.float64, device="cuda"); g = torch.empty_like(p)
    ms_g = timeit(lambda: m.log_prob_and_grad(p, lp, g))
    ms_l = timeit(lambda: m.log_prob(p, lp))
    rec = {"n": wl.n, "chains": C, "terms": wl.n_loc + wl.n_scale, "logp_grad_ms": ms_g, "logp_ms": ms_l,
           "logp_grad_evals_per_s": wl.n * C / ms_g * 1e3, "logp_evals_per_s": wl.n * C / ms_l * 1e3,
           "frac_of_hbm_roofline": wl.n * C / ms_g * 1e3 * wl.bytes_alg / 6560.3e9}
    if name.startswith("c4") or name.startswith("c3"):
        ms_x = timeit(lambda: m.xtwx(0, p))
        ms_p = timeit(lambda: m.precision(0, p))
        pdim = wl.model.all_terms()[0][0].basis.ncoef
        ms_all = timeit(lambda: m.precision_all(p))
        rec.update({"xtwx_term_ms": ms_x, "precision_chol_term_ms": ms_p, "precision_chol_all_loc_terms_ms": ms_all,
                    "n_loc_terms": wl.n_loc,
                    "xtwx_dense_equiv_tflops": 2.0 * wl.n * pdim * pdim * C / ms_x * 1e3 / 1e12,
                    "xtwx_banded_gflops": 2.0 * wl.n * 10 * C / ms_x * 1e3 / 1e9})
    if name.startswith("c3") or name.startswith("c2"):
        ms_w = timeit(lambda: m.pointwise(p))
        rec.update({"pointwise_waic_ms": ms_w, "pointwise_evals_per_s": wl.n * C / ms_w * 1e3})
    out[name] = rec
    print(name, json.dumps(rec))
json.dump(out, open("gpurun_out/r1_configs.json", "w"), indent=1)
