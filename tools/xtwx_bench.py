"""IWLS information sweep (precision + Cholesky of the location terms): float32 through the
tcgen05 route and the banded CUDA-core route, float64 through the banded route.
CUDA events, 3 warm-ups, mean of 5.  Writes gpurun_out/r1_xtwx.json."""
import json
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

import bench_workload as bw


def timeit(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


names = sys.argv[1:] or ["c3_n1M_5loc_5scale_nb20_1024chains", "c4_n1M_10loc_1scale_nb30_256chains"]
out = {}
for name in names:
    rec = {}
    for dt, tag in ((np.float32, "f32"), (np.float64, "f64")):
        wl = bw.build(name, dtype=dt)
        m = wl.model.build()
        p = torch.from_numpy(wl.params).cuda()
        for route in (("auto", "tcgen05", "tcgen05-w", "banded") if dt == np.float32 else ("banded",)):
            os.environ.pop("PTM_XTWX_TC", None)
            if route != "auto":
                os.environ["PTM_XTWX_TC"] = {"banded": "0", "tcgen05-w": "1", "tcgen05": "2"}[route]
            ms_all = timeit(lambda: m.precision_all(p))
            used_all = m.last_xtwx_route()
            ms_one = timeit(lambda: m.precision(0, p))
            used_one = m.last_xtwx_route()
            pdim = wl.model.all_terms()[0][0].basis.ncoef
            rec[f"{tag}_{route}"] = {
                "route_all": used_all, "route_one": used_one,
                "precision_chol_all_loc_terms_ms": ms_all, "precision_chol_one_term_ms": ms_one,
                "dense_equiv_tflops_all": 2.0 * wl.n * pdim * pdim * p.shape[0] * wl.n_loc / ms_all * 1e3 / 1e12,
            }
        os.environ.pop("PTM_XTWX_TC", None)
        del m, wl, p
        torch.cuda.empty_cache()
    out[name] = rec
    print(name, json.dumps(rec))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r1_xtwx.json", "w"), indent=1)
