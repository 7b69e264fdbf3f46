"""Bisects, through observation-shard views, to the observation on which the two float32 routes
(tcgen05 / CUDA cores) disagree most in the intercept gradients, and prints the oracle's view of
it (r, region code, grid cell).  Used for the one fuzz case whose gradient differed by 8.5e-4: a
single observation with r within 6e-8 of the core boundary a, where the declared h'' jumps between
the core lerp and the transition's closed form."""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests.helpers import build_case
from tests.test_gpu_parity import _f64_oracle_of
kw = dict(n=20011, C=128, n_loc=5, n_scale=7, nbases=30, dtype=np.float32, amp=0.15)
os.environ["PTM_MAIN_TC"] = "0"
cc = build_case(**kw)
os.environ.pop("PTM_MAIN_TC")
tc = build_case(**kw)
p = torch.from_numpy(cc.params).cuda()
def diff(lo, hi):
    b1 = tc.model.shard_view(lo, hi, add_priors=False).log_prob_and_grad_block(p).double()
    b0 = cc.model.shard_view(lo, hi, add_priors=False).log_prob_and_grad_block(p).double()
    d = (b1[:, 1:3] - b0[:, 1:3]).abs().amax(1)
    c = int(d.argmax())
    return float(d[c]), c, b1[c, :3].tolist(), b0[c, :3].tolist()
lo, hi = 0, 20011
for step in (1000, 100, 10, 1):
    best = (0, None)
    for a in range(lo, hi, step):
        e = diff(a, min(a + step, hi))
        if e[0] > best[0]: best = (e[0], a, e)
    print("step", step, "worst block", best[1], "err %.3e chain %d" % (best[2][0], best[2][1]), "tc", best[2][2], "cc", best[2][3], flush=True)
    lo, hi = best[1], min(best[1] + step, hi)
obs, chain = lo, best[2][1]
om, st = _f64_oracle_of(tc)
f = om.loglik_terms(st, chain)
print("obs", obs, "chain", chain, "r", f["r"][obs], "z", f["z"][obs], "d", f["d"][obs], "sd", f["sd"][obs], "code", int(om.spline.region_code(f["r"][obs:obs+1])[0]),
      "a,b", om.spline.min_knot, om.spline.max_knot, "min_eps,max_eps", om.spline.min_eps, om.spline.max_eps)
g = om.spline.bspline.grid
i = np.searchsorted(g, f["r"][obs], side="right") - 1
print("grid cell", i, "frac", (f["r"][obs] - g[i]) / om.spline.bspline.step)
