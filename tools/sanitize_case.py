"""Small shapes of every kernel family for compute-sanitizer (racecheck / memcheck):
  compute-sanitizer --tool racecheck python tools/sanitize_case.py
Covers the role-split shared-memory hand-offs SURVEY.md section 5 names as the risk area:
one term per term warp (TPW 1), two per warp (TPW 2, T > 12), shared designs, the logp-only
and pointwise modes, the banded and tcgen05 information sweeps, the predictive sweep."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import liesel_ptm_b200 as lp
from tests.helpers import build_case

cases = {
    "tpw1": dict(n=700, n_loc=2, n_scale=1, nbases=12, C=37),
    "tpw2": dict(n=500, n_loc=7, n_scale=7, nbases=10, C=33),
    "f32": dict(n=700, n_loc=2, n_scale=2, nbases=12, C=40, dtype=np.float32),
}
for name, kw in cases.items():
    case = build_case(**kw)
    p = torch.from_numpy(case.params).cuda()
    m = case.model
    lpv, g = m.log_prob_and_grad(p)
    m.log_prob(p)
    m.intercept_stats(p)
    m.pointwise(p)
    m.precision_all(p)
    m.precision(m.layout["beta"].__len__() - 1, p)
    m.summarise_dist(p[:5])
    v = m.shard_view(3, kw["n"] - 5, add_priors=False)
    v.log_prob_and_grad_block(p)
    torch.cuda.synchronize()
    print(name, "ok", float(lpv[0]), m.last_xtwx_route(), flush=True)

# shared designs: location and scale terms on the same covariates / knots
rng = np.random.default_rng(0)
n = 600
x = rng.uniform(-2, 2, size=(2, n))
y = rng.normal(size=n)
model = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=False)
for k in range(2):
    kn = lp.equidistant_knots(np.array([-2.0, 2.0]), 12)
    model.loc += lp.term.f_ig(lp.ps(x[k], nbases=12, xname=f"x{k}", knots=kn), fname="f")
    model.scale += lp.term.f_ig(lp.ps(x[k], nbases=12, xname=f"x{k}", knots=kn), fname="g")
model.build()
L, P = model.param_layout()
par = rng.normal(0, 0.05, size=(35, P))
par[:, L["tau"][0]:] = 1.0
lpv, g = model.log_prob_and_grad(torch.from_numpy(par).cuda())
torch.cuda.synchronize()
print("shared", model.shared_designs, "ok", float(lpv[0]), flush=True)
