"""Tensor-core route of the float sweep against the f64 oracle (small n) and the CUDA-core
route, then its time at the c3 shape."""
import os, sys, json
sys.path.insert(0, '.')
import numpy as np, torch
from tests.helpers import build_case, oracle_eval, rel_err
import bench_workload as bw

def t_ms(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for shape in (dict(n=4097, n_loc=2, n_scale=2, nbases=12, C=70), dict(n=20011, n_loc=5, n_scale=5, nbases=20, C=130),
              dict(n=333, n_loc=1, n_scale=0, nbases=9, C=3), dict(n=50, n_loc=0, n_scale=1, nbases=20, C=257)):
    case = build_case(dtype=np.float32, **shape)
    p = torch.from_numpy(case.params).cuda()
    lp, g = case.model.log_prob_and_grad(p)
    lp_only = case.model.log_prob(p)
    torch.cuda.synchronize()
    route = case.model.last_main_route()
    case64 = build_case(dtype=np.float32, build=False, **shape)
    lp_ref, g_ref = oracle_eval(case64)   # float32 oracle arithmetic; compare loosely, then to the other route
    os.environ["PTM_MAIN_TC"] = "0"
    c2 = build_case(dtype=np.float32, **shape)
    os.environ.pop("PTM_MAIN_TC")
    lp0, g0 = c2.model.log_prob_and_grad(p)
    st0 = c2.model.intercept_stats(p); st1 = case.model.intercept_stats(p)
    print(shape, route, c2.model.last_main_route(),
          "logp vs cuda-core %.2e  grad %.2e  logp-only %.2e  stats %.2e | vs f32 oracle logp %.2e grad %.2e" % (
          rel_err(lp.cpu().numpy(), lp0.cpu().numpy()), max(rel_err(g[c].cpu().numpy(), g0[c].cpu().numpy()) for c in range(g.shape[0])),
          rel_err(lp_only.cpu().numpy(), lp0.cpu().numpy()), rel_err(st1.cpu().numpy(), st0.cpu().numpy()),
          rel_err(lp.cpu().numpy(), lp_ref), rel_err(g.cpu().numpy(), g_ref)), flush=True)

wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", dtype=np.float32)
m = wl.model.build()
p = torch.from_numpy(wl.params).cuda()
lpv = torch.empty(p.shape[0], dtype=torch.float32, device="cuda"); g = torch.empty_like(p)
a = t_ms(lambda: m.log_prob_and_grad(p, lpv, g)); r = m.last_main_route()
b = t_ms(lambda: m.log_prob(p, lpv))
m.profile(True); m.log_prob_and_grad(p, lpv, g); k = m.last_kernel_ms(); m.profile(False)
print(f"c3 f32 {r}: logp+grad {a:.2f} ms, logp {b:.2f} ms, kernels {k}")
lp1, g1 = lpv.clone(), g.clone()
os.environ["PTM_MAIN_TC"] = "0"
wl2 = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", dtype=np.float32)
m2 = wl2.model.build()
lp0, g0 = m2.log_prob_and_grad(p)
print("c3 tc vs cuda-core: logp %.2e grad %.2e" % (rel_err(lp1.cpu().numpy(), lp0.cpu().numpy()),
      max(rel_err(g1[c].cpu().numpy(), g0[c].cpu().numpy()) for c in range(0, 1024, 37))))

# accuracy at n = 1M against the float64 kernel on the float32-rounded inputs (64 chains)
import liesel_ptm_b200 as lp_
def against_f64(tag):
    C = 64
    w32 = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=C, dtype=np.float32)
    m32 = w32.model.build()
    m64 = lp_.LocScalePTM(w32.y.astype(np.float64), m32.knots.astype(np.float64), to_float32=False)
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
    lp32, g32 = m32.log_prob_and_grad(p32); route = m32.last_main_route()
    lp64, g64 = m64.log_prob_and_grad(p32.double())
    e_lp = ((lp32.double() - lp64).abs() / lp64.abs()).max().item()
    e_g = ((g32.double() - g64).abs().amax(dim=1) / g64.abs().amax(dim=1))
    print(f"[{tag} {route}] n=1M vs f64 kernel: logp {e_lp:.2e}  grad max {e_g.max().item():.2e} median {e_g.median().item():.2e}")
os.environ.pop("PTM_MAIN_TC", None)
against_f64("tc")
os.environ["PTM_MAIN_TC"] = "0"
against_f64("cuda-core")
