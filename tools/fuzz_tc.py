"""Random-shape sweep of the float32 tensor-core route against the float32 CUDA-core route
(PTM_MAIN_TC=0 at create) and the float64 oracle on the float32-rounded inputs: logp, gradient,
logp-only sweep, all three transformation variants.  python tools/fuzz_tc.py [n_cases] [seed]"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests.helpers import build_case, pack_grad_layout, rel_err
from tests.test_gpu_parity import _f64_oracle_of

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for k in range(ncase):
    n = int(rng.choice([1, 31, 47, 48, 49, 95, 97, 191, 193, 777, 4097, 20011]))
    C = int(rng.choice([1, 2, 33, 127, 128, 129, 300]))
    n_loc, n_scale = int(rng.integers(0, 8)), int(rng.integers(0, 8))
    if rng.uniform() < 0.25:  # many terms: several-pass forms, two terms per stager warp
        n_loc, n_scale = int(rng.integers(6, 13)), int(rng.integers(5, 12))
    if n_loc + n_scale == 0:
        n_loc = 1
    nb = int(rng.choice([6, 8, 12, 20, 21, 30]))
    variant = str(rng.choice(["ptm", "ptm", "onion", "identity"]))
    kw = dict(n=n, C=C, n_loc=n_loc, n_scale=n_scale, nbases=nb, dtype=np.float32, amp=0.15, bspline=variant)
    os.environ.pop("PTM_MAIN_TC", None)
    tc = build_case(**kw)
    os.environ["PTM_MAIN_TC"] = "0"
    cc = build_case(**kw)
    os.environ.pop("PTM_MAIN_TC", None)
    p = torch.from_numpy(tc.params).cuda()
    try:
        lp, g = tc.model.log_prob_and_grad(p)
    except Exception as exc:  # a documented capacity limit is a loud error, not a wrong result
        print(f"[{k:2d}] n={n} C={C} loc={n_loc} scale={n_scale} nb={nb} {variant}  REFUSED: {exc}")
        continue
    route = tc.model.last_main_route()
    lpo = tc.model.log_prob(p)
    lp0, g0 = cc.model.log_prob_and_grad(p)
    assert cc.model.last_main_route() == "cuda-core"
    e_lp = max(rel_err(lp.cpu().numpy(), lp0.cpu().numpy()), rel_err(lpo.cpu().numpy(), lp0.cpu().numpy()))
    e_g = float(((g.double() - g0.double()).abs().amax(1) / g0.double().abs().amax(1).clamp_min(1e-30)).max())
    tag = f"n={n} C={C} loc={n_loc} scale={n_scale} nb={nb} {variant}"
    extra = ""
    ok = e_lp < 5e-6 and e_g < 3e-5
    if n <= 4097 and C <= 33 and variant == "ptm":
        om, st = _f64_oracle_of(tc)
        lp_ref, gd = om.logp_and_grad(st)
        g_ref = pack_grad_layout(tc.layout, tc.P, gd)
        e_o = max(rel_err(lp.cpu().numpy(), lp_ref), max(rel_err(g[c].cpu().numpy(), g_ref[c]) for c in range(C)))
        extra = f" vs f64 oracle {e_o:.1e}"
        ok = ok and e_o < 1e-5
    bad += 0 if ok else 1
    print(f"[{k:2d}] {tag:50s} {route:9s} logp {e_lp:.1e} grad {e_g:.1e}{extra}  {'ok' if ok else 'FAIL'}", flush=True)
    del tc, cc
    torch.cuda.empty_cache()
print("failures:", bad)
