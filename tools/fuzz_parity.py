"""Random-shape parity sweep (GPU vs oracle): logp, gradient, information matrices, pointwise
records, both precisions.  python tools/fuzz_parity.py [n_cases] [seed]"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from tests.helpers import build_case, oracle_eval, rel_err

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
worst = {}
for k in range(ncase):
    n = int(rng.choice([1, 2, 3, 17, 40, 41, 79, 80, 81, 333, 1000, 2049, 7001]))
    C = int(rng.choice([1, 2, 31, 32, 33, 64, 65, 129]))
    n_loc, n_scale = int(rng.integers(0, 5)), int(rng.integers(0, 4))
    if rng.uniform() < 0.3:  # more than 12 terms: two terms per term warp
        n_loc, n_scale = int(rng.integers(6, 13)), int(rng.integers(5, 13))
    nb = int(rng.choice([6, 8, 12, 20, 21, 30]))
    dt = np.float64 if rng.uniform() < 0.7 else np.float32
    case = build_case(n=n, C=C, n_loc=n_loc, n_scale=n_scale, nbases=nb, dtype=dt)
    p = torch.from_numpy(case.params).cuda()
    try:
        lp, g = case.model.log_prob_and_grad(p)
    except Exception as exc:  # a documented capacity limit is a loud error, not a wrong result
        print(f"[{k:2d}] n={n} C={C} loc={n_loc} scale={n_scale} nb={nb} {np.dtype(dt).name}  REFUSED: {exc}")
        continue
    lp, g = lp.cpu().numpy().astype(np.float64), g.cpu().numpy().astype(np.float64)
    lp2 = case.model.log_prob(p).cpu().numpy().astype(np.float64)
    lp_ref, g_ref = oracle_eval(case)
    tol = 1e-11 if dt == np.float64 else 3e-4
    e = max(rel_err(lp, lp_ref), rel_err(lp2, lp_ref), rel_err(g, g_ref))
    tag = f"n={n} C={C} loc={n_loc} scale={n_scale} nb={nb} {np.dtype(dt).name}"
    status = "ok" if e < tol else "FAIL"
    extra = ""
    if dt == np.float64 and n_loc > 0:
        P, L = case.model.precision(0, p)
        _, rP, rL = case.omodel.loc_precision(case.state, 0)
        e2 = rel_err(P.cpu().numpy(), rP)
        extra = f" prec {e2:.1e}"
        if e2 > 1e-10: status = "FAIL"
    if dt == np.float32 and n_loc > 0:
        # information sweeps: both tensor-core versions against the banded CUDA-core route
        import os
        mats = {}
        for route in ("0", "1", "2"):
            os.environ["PTM_XTWX_TC"] = route
            Ps, _ = case.model.precision_all(p)
            mats[route] = (np.stack([x.cpu().numpy() for x in Ps]).astype(np.float64), case.model.last_xtwx_route())
        os.environ.pop("PTM_XTWX_TC", None)
        e4 = max(rel_err(mats["1"][0], mats["0"][0]), rel_err(mats["2"][0], mats["0"][0]))
        extra += f" info[{mats['1'][1]},{mats['2'][1]}] {e4:.1e}"
        if not e4 < 5e-5: status = "FAIL"
    if dt == np.float64:
        lppd, pw = case.model.pointwise(p)
        ll = np.stack([case.omodel.loglik_terms(case.state, c)["ll"] for c in range(C)])
        m = ll.max(0); ref = m + np.log(np.exp(ll - m).sum(0)) - np.log(C)
        e3 = rel_err(lppd.cpu().numpy(), ref)
        extra += f" lppd {e3:.1e}"
        if e3 > 1e-11: status = "FAIL"
    print(f"[{k:2d}] {tag:55s} logp/grad {e:.1e}{extra}  {status}", flush=True)
    worst[status] = worst.get(status, 0) + 1
print(worst)
sys.exit(1 if worst.get("FAIL") else 0)
