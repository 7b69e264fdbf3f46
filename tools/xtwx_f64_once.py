import sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench_workload as bw
def timeit(fn, reps=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name in ("c3_n1M_5loc_5scale_nb20_1024chains", "c4_n1M_10loc_1scale_nb30_256chains"):
    wl = bw.build(name, dtype=np.float64)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).cuda()
    print(name[:2], "f64 precision_all %.2f ms" % timeit(lambda: m.precision_all(p)), m.last_xtwx_route(), flush=True)
    del m, wl, p; torch.cuda.empty_cache()
