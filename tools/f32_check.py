"""f32 kernel vs f64 kernel on the same f32-rounded inputs at a given size (GPU only)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
name = "c3_n1M_5loc_5scale_nb20_1024chains"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
w32 = bw.build(name, chains=C, dtype=np.float32, n_override=n)
m32 = w32.model.build()
# f64 model on the f32-rounded numbers
import liesel_ptm_b200 as lp
m64 = lp.LocScalePTM(w32.y.astype(np.float64), m32.knots.astype(np.float64), to_float32=False)
for t, k in w32.model.all_terms():
    b = t.basis
    b64 = lp.Basis(x=b.x.astype(np.float64), xname=b.xname, knots=b.knots.astype(np.float64),
                   Z=b.Z.astype(np.float64), penalty=b.penalty.astype(np.float64), nbases=b.nbases)
    tm = lp.term.f_ig(b64, fname="f" if k == 0 else "g")
    tm.penalty_rank = t.penalty_rank
    if k == 0: m64.loc += tm
    else: m64.scale += tm
m64.build()
# NOTE: trafo_Z differs in rounding between the two builds only at the 1e-8 level (f32 rounding of Z)
p32 = torch.from_numpy(w32.params).cuda(); p64 = p32.double()
lp32, g32 = m32.log_prob_and_grad(p32); lp64, g64 = m64.log_prob_and_grad(p64)
torch.cuda.synchronize()
e_lp = ((lp32.double() - lp64).abs() / lp64.abs()).max().item()
e_g = ((g32.double() - g64).abs().amax(dim=1) / g64.abs().amax(dim=1)).max().item()
print(f"n={n} C={C}: f32 vs f64 kernel: logp rel {e_lp:.3e}  grad rel(inf-norm per chain) {e_g:.3e}")
for _ in range(3): m32.log_prob_and_grad(p32)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): m32.log_prob_and_grad(p32)
torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 5 * 1e3
print(f"f32 logp+grad {ms:.2f} ms -> {n*C/ms*1e3:.3e} evals/s")
L, P = m32.param_layout()
d = (g32.double() - g64).abs()
scale = g64.abs().amax(dim=1, keepdim=True)
rel = (d / scale)
blocks = {"beta0": (0, 1), "gamma0": (1, 2), "delta_z": (L["delta_z"], L["delta_z"] + 19), "tau": (L["tau"][0], L["tau"][0] + 10), "tau_delta": (L["tau_delta"], L["tau_delta"] + 1)}
for t in range(10): blocks[f"beta{t}"] = (L["beta"][t], L["beta"][t] + 19)
for k, (a_, b_) in blocks.items():
    print(f"  {k:10s} max rel-to-inf-norm err {rel[:, a_:b_].max().item():.2e}   |g| max {g64[:, a_:b_].abs().max().item():.3e}")
