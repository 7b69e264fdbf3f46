"""Accuracy of the float sweep routes at n = 1M against the float64 kernel on the float32-rounded
inputs (64 chains of c3)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
import liesel_ptm_b200 as lp_
C = 64
w32 = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=C, dtype=np.float32)
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
L, P = m64.param_layout()
for env in sys.argv[1:] or [""]:
    for kv in env.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    w = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=C, dtype=np.float32)
    m32 = w.model.build()
    lp32, g32 = m32.log_prob_and_grad(p32); route = m32.last_main_route()
    e_lp = ((lp32.double() - lp64).abs() / lp64.abs()).max().item()
    d = (g32.double() - g64).abs()
    e_g = (d.amax(dim=1) / g64.abs().amax(dim=1))
    nb = L["delta_z"]
    e_beta = (d[:, 2:nb].amax(dim=1) / g64.abs().amax(dim=1)).max().item()
    e_dz = (d[:, nb:nb + 19].amax(dim=1) / g64.abs().amax(dim=1)).max().item()
    e_ic = (d[:, :2].amax(dim=1) / g64.abs().amax(dim=1)).max().item()
    print(f"[{env or 'default'} -> {route}] logp {e_lp:.2e}  grad max {e_g.max().item():.2e} median {e_g.median().item():.2e}"
          f"  (beta block {e_beta:.2e}, delta_z block {e_dz:.2e}, intercepts {e_ic:.2e})", flush=True)
    for kv in env.split(","):
        if "=" in kv: os.environ.pop(kv.split("=")[0])
