import sys
sys.path.insert(0, '.')
import numpy as np, torch
import bench_workload as bw
dt = np.float32 if (len(sys.argv) < 2 or sys.argv[1] == "f32") else np.float64
wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", dtype=dt)
m = wl.model.build()
p = torch.from_numpy(wl.params).cuda()
for _ in range(2):
    lp, g = m.log_prob_and_grad(p)
torch.cuda.synchronize()
print(m.last_main_route(), float(lp[0]))
