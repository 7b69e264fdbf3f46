"""Two float32 precision_all calls on the c3 workload (profiling target for k_xtwx_tc)."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
import bench_workload as bw
name = sys.argv[1] if len(sys.argv) > 1 else "c3_n1M_5loc_5scale_nb20_1024chains"
wl = bw.build(name, dtype=np.float32)
m = wl.model.build()
p = torch.from_numpy(wl.params).cuda()
for _ in range(2):
    m.precision_all(p)
torch.cuda.synchronize()
print("route", m.last_xtwx_route())
