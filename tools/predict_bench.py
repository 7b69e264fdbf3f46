"""Timing of the predictive sweeps and the inverse (run under gpurun):
python tools/predict_bench.py [m] [S] [f32]  ->  gpurun_out/r1_predict.json"""
import json, os, sys
sys.path.insert(0, '.')
import numpy as np, torch
import liesel_ptm_b200 as lp

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
F32 = len(sys.argv) > 3 and sys.argv[3] == 'f32'
DT = np.float32 if F32 else np.float64
w = 4 if F32 else 8
rng = np.random.default_rng(0)
T, n = 10, 100_000
X = rng.uniform(-2, 2, size=(T, n)).astype(DT)
y = (rng.normal(size=n) * 1.5).astype(DT)
model = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=F32)
for t in range(T):
    tm = lp.term.f_ig(lp.ps(X[t], nbases=20, xname=f"x{t}", dtype=DT))
    if t < 5: model.loc += tm
    else: model.scale += tm
model.build()
L, P = model.param_layout()
par = np.zeros((S, P)); par[:, 2:L['delta_z']] = rng.normal(size=(S, L['delta_z'] - 2)) * 0.05
par[:, L['delta_z']:L['delta_z'] + 19] = rng.normal(size=(S, 19)) * 0.2
par[:, L['tau'][0]:] = 1.0
p = torch.from_numpy(par.astype(DT)).cuda()
lo = max(X[t].min() for t in range(T)); hi = min(X[t].max() for t in range(T))
newdata = {f"x{t}": rng.uniform(lo, hi, size=m).astype(DT) for t in range(T)}
grid = (rng.normal(size=m) * 1.5).astype(DT)
u = rng.uniform(0.001, 0.999, size=m).astype(DT)
g, Xd, _ = model._newdata(newdata, grid)
ud = torch.from_numpy(u).cuda()
ws = torch.empty(int(model._lib.ptm_predict_workspace_bytes(model._handle, S)), dtype=torch.uint8, device='cuda')
out = [torch.empty((S, m), dtype=p.dtype, device='cuda') for _ in range(3)]
sfx = 'f32' if F32 else 'f64'
fp = getattr(model._lib, f"ptm_predict_{sfx}"); fq = getattr(model._lib, f"ptm_predict_quantile_{sfx}")

def timed(fn, K=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K

st = model._stream()
res = {"m": m, "S": S, "dtype": sfx, "T": T}
ms = timed(lambda: fp(model._handle, p.data_ptr(), S, g.data_ptr(), Xd.data_ptr(), m, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), ws.data_ptr(), ws.numel(), st))
res["predict_z_logpdf_cdf_ms"] = ms; res["predict_all_write_GBps"] = 3 * S * m * w / ms / 1e6
ms = timed(lambda: fp(model._handle, p.data_ptr(), S, g.data_ptr(), Xd.data_ptr(), m, 0, out[1].data_ptr(), 0, ws.data_ptr(), ws.numel(), st))
res["predict_logpdf_ms"] = ms; res["predict_logpdf_write_GBps"] = S * m * w / ms / 1e6
ms = timed(lambda: fq(model._handle, p.data_ptr(), S, ud.data_ptr(), Xd.data_ptr(), m, out[0].data_ptr(), ws.data_ptr(), ws.numel(), st))
res["quantile_ms"] = ms; res["quantile_write_GBps"] = S * m * w / ms / 1e6
delta = torch.from_numpy((rng.normal(size=(S, 20)) * 0.3).astype(DT)).cuda()
z = torch.from_numpy((rng.normal(size=min(m, 1_000_000)) * 2).astype(DT)).cuda()
ms = timed(lambda: model.inverse(delta, z))
res["inverse_ms"] = ms; res["inverse_evals_per_s"] = S * z.numel() / ms * 1e3
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(f"gpurun_out/r1_predict_{sfx}.json", "w"), indent=1)
