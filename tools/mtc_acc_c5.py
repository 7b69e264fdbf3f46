"""Both float32 routes against the float64 kernel on the float32-rounded inputs for the per-GPU
slice of config 5 (n = 2M, 10 + 10 terms, 64 chains), with the half-ulp conditioning kappa."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
import liesel_ptm_b200 as lp_
C = 64
name = "c5_n50M_10loc_10scale_128chains"
w32 = bw.build(name, chains=C, dtype=np.float32, n_override=2_000_000)
m64 = lp_.LocScalePTM(w32.y.astype(np.float64), w32.model.knots.astype(np.float64), to_float32=False)
for t, k in w32.model.all_terms():
    b = t.basis
    tm = lp_.term.f_ig(lp_.Basis(x=b.x.astype(np.float64), xname=b.xname, knots=b.knots.astype(np.float64),
                                  Z=b.Z.astype(np.float64), penalty=b.penalty.astype(np.float64), nbases=b.nbases),
                       fname="f" if k == 0 else "g")
    tm.penalty_rank = t.penalty_rank
    if k == 0: m64.loc += tm
    else: m64.scale += tm
m64.build()
p32 = torch.from_numpy(w32.params).cuda()
lp64, g64 = m64.log_prob_and_grad(p32.double())
gen = torch.Generator(device="cuda"); gen.manual_seed(5)
half_ulp = (torch.rand(p32.shape, dtype=torch.float64, device="cuda", generator=gen) - 0.5) * 2.0 ** -23
_, g64p = m64.log_prob_and_grad(p32.double() * (1.0 + half_ulp))
kappa = ((g64p - g64).abs().amax(dim=1) / g64.abs().amax(dim=1)).max().item()
print(f"half-ulp conditioning kappa {kappa:.2e}")
for env in ("", "PTM_MAIN_TC=0"):
    if env: os.environ["PTM_MAIN_TC"] = "0"
    w = bw.build(name, chains=C, dtype=np.float32, n_override=2_000_000)
    m32 = w.model.build()
    lp32, g32 = m32.log_prob_and_grad(p32)
    e_lp = ((lp32.double() - lp64).abs() / lp64.abs()).max().item()
    e_g = ((g32.double() - g64).abs().amax(dim=1) / g64.abs().amax(dim=1))
    print(f"[{env or 'default'} -> {m32.last_main_route()}] logp {e_lp:.2e}  grad max {e_g.max().item():.2e} median {e_g.median().item():.2e}", flush=True)
    os.environ.pop("PTM_MAIN_TC", None)
