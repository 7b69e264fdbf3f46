"""Stress checks outside the test suite: 5000 chains against the oracle; n = 3e7 (> 2^24 observations)
with ragged shards adding up to the full evaluation (64-bit observation indexing)."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch, time
import bench_workload as bw
from tests.helpers import build_case, oracle_eval, rel_err
import liesel_ptm_b200 as lp
# (a) many chains
case = build_case(n=1500, n_loc=2, n_scale=1, nbases=12, C=5000)
p = torch.from_numpy(case.params).cuda()
lpv, g = case.model.log_prob_and_grad(p)
lp_ref, g_ref = oracle_eval(case)
print("C=5000:", rel_err(lpv.cpu().numpy(), lp_ref), rel_err(g.cpu().numpy(), g_ref))
P, L = case.model.precision(0, p)
_, rP, rL = case.omodel.loc_precision(case.state, 0)
print("C=5000 precision:", rel_err(P.cpu().numpy(), rP))
# (b) many observations: n = 3e7 > 2^24, shards add up, last observations matter
n = 30_000_000
y, X = bw.make_data(n, 1, 1, seed=3)
m = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=False)
kn = lp.equidistant_knots(np.array([-2.0, 2.0]), 20)
sub = slice(0, 200000)
for t in range(2):
    b0 = lp.ps(X[t][sub], nbases=20, xname=f"x{t}", knots=kn)
    b = lp.Basis(x=np.clip(X[t], -2 + 1e-9, 2 - 1e-9), xname=f"x{t}", knots=b0.knots, Z=b0.Z, penalty=b0.penalty, nbases=20)
    (m.loc if t == 0 else m.scale).__iadd__(lp.term.f_ig(b, fname="f" if t == 0 else "g"))
m.build()
par = torch.from_numpy(bw.make_states(m, 40, seed=2)).cuda()
full, gfull = m.log_prob_and_grad(par); full, gfull = full.clone(), gfull.clone()
acc = torch.zeros_like(full); gacc = torch.zeros_like(gfull)
for lo, hi, pri in ((0, 1, True), (1, 16_777_217, False), (16_777_217, n - 1, False), (n - 1, n, False)):
    m.set_shard(lo, hi, add_priors=pri)
    a, b_ = m.log_prob_and_grad(par); acc += a; gacc += b_
print("n=3e7 shards:", float(((acc - full).abs() / full.abs()).max()), float((gacc - gfull).abs().max() / gfull.abs().max()))
