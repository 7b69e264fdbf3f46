"""ESS/s of the repo's minimal sampler (IWLS-MH per smooth term + HMC on delta_z, smoothing
scales fixed) over the fused operator, for BASELINE configs 1 and 2.  min bulk-ESS over the
sampled parameters / wall seconds of the posterior epoch."""
import json, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
from liesel_ptm_b200 import mcmc

out = {}
for name, C, warm, draws in [("c1_readme_n100", 64, 300, 500), ("c2_n100k_3loc_1scale_64chains", 64, 200, 300)]:
    wl = bw.build(name, chains=C, dtype=np.float64)
    m = wl.model.build()
    start = wl.params.copy()
    L0 = m.layout
    start[:, 2:L0["delta_z"] + m.nparam - 1] *= 0.01  # start near zero like the reference's initial values
    p0 = torch.from_numpy(start).cuda()
    dr, info = mcmc.run_chains(m, p0, warmup=warm, draws=draws, seed=0)
    L = m.layout
    cols = list(range(L["beta"][0], L["delta_z"] + m.nparam - 1)) if m.terms else list(range(L["delta_z"], L["delta_z"] + m.nparam - 1))
    ess = np.array([mcmc.effective_sample_size(dr[:, :, j].T) for j in cols])
    print("ess per parameter", np.round(ess).astype(int).tolist())
    rec = {"chains": C, "warmup": warm, "draws": draws, "posterior_seconds": info["posterior_seconds"],
           "min_ess": float(ess.min()), "median_ess": float(np.median(ess)),
           "min_ess_per_s": float(ess.min() / info["posterior_seconds"]),
           "iterations_per_s": draws / info["posterior_seconds"],
           "accept_iwls": info["accept_iwls"], "accept_dz": info["accept_dz"], "n_leapfrog": info["n_leapfrog"], "cuda_graph": info["cuda_graph"]}
    out[name] = rec
    print(name, json.dumps(rec))
json.dump(out, open("gpurun_out/r1_ess.json", "w"), indent=1)
