"""Secondary measurements bench.py attaches under `extra` (same run, same box): the float32
c3 line, the c4 information sweep, ESS/s of the minimal sampler on c1/c2 (N = 1); strong
scaling of c3 and the observation-sharded c5 step with its all-reduce timed apart (N > 1).
Every entry is measured with CUDA events on the launching stream after warm-up; a failing
entry reports its error text instead of killing the headline line."""

from __future__ import annotations

import time
import traceback

import numpy as np


def _events(torch, fn, reps, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _guard(out, key, fn):
    t0 = time.perf_counter()
    try:
        out[key] = fn()
    except Exception as e:  # report, do not hide
        out[key] = {"error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc(limit=3)}
    if isinstance(out[key], dict):
        out[key]["wall_s"] = round(time.perf_counter() - t0, 2)


def c3_other_dtype(torch, bw, dev, peak_gbs, reps, dtype):
    """The headline workload at the other precision (float64 when the headline is float32)."""
    wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", dtype=dtype, device=dev)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).to(dev)
    lp = torch.empty(p.shape[0], dtype=p.dtype, device=dev)
    g = torch.empty_like(p)
    ms = _events(torch, lambda: m.log_prob_and_grad(p, lp, g), reps, 3)
    m.profile(True)
    m.log_prob_and_grad(p, lp, g)
    k_ms = m.last_kernel_ms()[1]
    m.profile(False)
    evals = wl.n * p.shape[0]
    ach = evals * wl.bytes_alg / (k_ms * 1e-3) / 1e9
    return {"workload": wl.name, "dtype": "f64" if np.dtype(dtype) == np.float64 else "f32", "route": m.last_main_route(),
            "ms_per_step": ms, "kernel_ms": k_ms,
            "value": evals / (ms * 1e-3), "unit": "obs*chain evals/s", "bytes_alg_per_eval": wl.bytes_alg,
            "roofline_frac": ach / peak_gbs, "finite": bool(torch.isfinite(lp).all() and torch.isfinite(g).all())}


def c4_information(torch, bw, dev, reps):
    out = {}
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        wl = bw.build("c4_n1M_10loc_1scale_nb30_256chains", dtype=dt, device=dev)
        m = wl.model.build()
        p = torch.from_numpy(wl.params).to(dev)
        ms = _events(torch, lambda: m.precision_all(p), reps, 2)
        route = m.last_xtwx_route()
        pdim = wl.model.all_terms()[0][0].basis.ncoef
        out[tag] = {"precision_chol_all_loc_terms_ms": ms, "route": route,
                    "dense_equiv_tflops": 2.0 * wl.n * pdim * pdim * p.shape[0] * wl.n_loc / (ms * 1e-3) / 1e12,
                    "chains": int(p.shape[0]), "p": int(pdim), "terms": wl.n_loc}
        ms_lg = _events(torch, lambda: m.log_prob_and_grad(p), reps, 2)
        out[tag]["logp_grad_ms"] = ms_lg
        out[tag]["logp_grad_evals_per_s"] = wl.n * p.shape[0] / (ms_lg * 1e-3)
        del m, wl, p
        torch.cuda.empty_cache()
    out["flops_definition"] = "2 n p^2 per chain*term (Z.T @ (Z*w), iwls_proposals.py:69-74); the sweep itself uses the banded identity / tcgen05 transposed-band route"
    return out


def ess(torch, bw, dev):
    from liesel_ptm_b200 import mcmc

    out = {}
    for name, C, warm, draws in (("c1_readme_n100", 64, 300, 500), ("c2_n100k_3loc_1scale_64chains", 64, 200, 300)):
        wl = bw.build(name, chains=C, dtype=np.float64, device=dev)
        m = wl.model.build()
        start = wl.params.copy()
        L0 = m.layout
        start[:, 2:L0["delta_z"] + m.nparam - 1] *= 0.01
        rec = {}
        for tag, kw in (("fixed_tau", {}), ("reference_composition", dict(sample_tau=True, pgibbs_intercepts=True))):
            dr, info = mcmc.run_chains(m, torch.from_numpy(start).to(dev), warmup=warm, draws=draws, seed=0, **kw)
            cols = list(range(L0["beta"][0] if m.terms else L0["delta_z"], L0["delta_z"] + m.nparam - 1))
            if kw:
                cols = cols + [L0["beta0"], L0["gamma0"]] + list(L0["tau"])
            e = np.array([mcmc.effective_sample_size(dr[:, :, j].T) for j in cols])
            rec[tag] = {"posterior_seconds": info["posterior_seconds"], "min_ess": float(e.min()),
                        "min_ess_per_s": float(e.min() / info["posterior_seconds"]),
                        "iterations_per_s": draws / info["posterior_seconds"], "cuda_graph": info["cuda_graph"],
                        "accept_iwls": info["accept_iwls"], "accept_dz": info["accept_dz"]}
        out[name] = {"chains": C, "warmup": warm, "draws": draws, **rec,
                     "sampler": "liesel_ptm_b200/mcmc.py over the fused operator.  fixed_tau: IWLS-MH per smooth term + HMC "
                                "on delta_z.  reference_composition: + inverse-gamma Gibbs on every tau^2 (gam/kernel.py:9-41) "
                                "+ pseudo-Gibbs intercepts (predictor.py:133-175, 239-281); delta_z by HMC (NUTS in the "
                                "reference), tau_delta fixed; min ESS over coefficients, intercepts and scales"}
    return out


def strong_scaling_c3(torch, dist, bw, dev, rank, world, dtype, reps):
    """The 1024 chains of c3 split over the ranks (1024 / N per GPU): total work fixed."""
    wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", dtype=dtype, device=dev)
    C_total = wl.params.shape[0]
    lo, hi = rank * C_total // world, (rank + 1) * C_total // world
    m = wl.model.build()
    p = torch.from_numpy(wl.params[lo:hi]).to(dev)
    lp = torch.empty(p.shape[0], dtype=p.dtype, device=dev)
    g = torch.empty_like(p)
    for _ in range(3):
        m.log_prob_and_grad(p, lp, g)
    torch.cuda.synchronize()
    dist.barrier()
    ms = _events(torch, lambda: m.log_prob_and_grad(p, lp, g), reps, 0)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return {"workload": wl.name, "scaling": "strong", "chains_total": int(C_total), "chains_per_gpu": int(hi - lo),
            "ms_per_step": ms, "value": wl.n * C_total / (ms * 1e-3), "unit": "obs*chain evals/s"}


def c5_obs_sharded(torch, dist, bw, lp_mod, dev, rank, world, dtype, reps, n_per_gpu=6_250_000):
    """BASELINE config 5 shape: observations sharded over the ranks (6.25M per GPU: 50M at
    N = 8), 10 + 10 smooth terms, 128 chains on every rank, ONE NCCL all-reduce of the packed
    [C, 1 + P] block the epilogue kernel writes in place."""
    from liesel_ptm_b200 import parallel

    n_loc = n_scale = 10
    T, nbases, C = 20, 20, 128
    y_sub, X_sub = bw.make_data(200_000, n_loc, n_scale, seed=1234)
    bases = [lp_mod.ps(X_sub[t].astype(dtype), nbases=nbases, xname=f"x{t}", dtype=dtype,
                       knots=lp_mod.equidistant_knots(np.array([-2.0, 2.0]), nbases, dtype=dtype)) for t in range(T)]
    y, X = bw.make_data(n_per_gpu, n_loc, n_scale, seed=100 + rank)
    X = np.clip(X, -2.0 + 1e-9, 2.0 - 1e-9)
    f32 = np.dtype(dtype) == np.float32
    model = lp_mod.LocScalePTM.new_ptm(y.astype(dtype), nparam=20, to_float32=f32, device=dev)
    sub = lp_mod.LocScalePTM.new_ptm(y_sub.astype(dtype), nparam=20, to_float32=f32, device=dev)
    for t in range(T):
        b = lp_mod.Basis(x=X[t].astype(dtype), xname=f"x{t}", knots=bases[t].knots, Z=bases[t].Z,
                         penalty=bases[t].penalty, nbases=nbases)
        for mdl, bb in ((model, b), (sub, bases[t])):
            tm = lp_mod.term.f_ig(bb, fname="f" if t < n_loc else "g")
            if t < n_loc:
                mdl.loc += tm
            else:
                mdl.scale += tm
    model.build()
    del X
    view = model.shard_view(0, n_per_gpu, add_priors=(rank == 0))
    params = torch.from_numpy(bw.make_states(sub, C, seed=7).astype(dtype)).to(dev)
    block = torch.empty((C, 1 + params.shape[1]), dtype=params.dtype, device=dev)

    def local():
        view.log_prob_and_grad_block(params, block)

    def step():
        local()
        parallel.allreduce_block_(block)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    ms = _events(torch, step, reps, 0)
    dist.barrier()
    ms_local = _events(torch, local, reps, 0)
    dist.barrier()
    ms_ar = _events(torch, lambda: parallel.allreduce_block_(block), 20, 3)
    t = torch.tensor([ms, ms_local, ms_ar], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_local, ms_ar = (float(v) for v in t.tolist())
    # the reduced logp equals the sum of the partials gathered one by one
    local()
    part = block[:, 0].clone()
    gathered = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    ref = torch.stack(gathered).sum(0)
    step()
    err = float(((block[:, 0] - ref).abs() / ref.abs().clamp_min(1.0)).max())
    n_total = n_per_gpu * world
    return {"workload": "c5 shape: observation-sharded, 10 + 10 terms nbases 20, 128 chains", "scaling": "weak in n",
            "n_total": n_total, "n_per_gpu": n_per_gpu, "chains": C, "params_per_chain": int(params.shape[1]),
            "ms_per_step": ms, "local_sweep_ms": ms_local, "allreduce_ms": ms_ar,
            "allreduce_bytes": int(block.numel() * block.element_size()),
            "value": n_total * C / (ms * 1e-3), "unit": "obs*chain evals/s",
            "allreduce_check_rel_err": err, "finite": bool(torch.isfinite(block).all())}
