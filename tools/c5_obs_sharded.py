"""BASELINE config 5: observation-sharded PTM, n = 50M, 20 smooth terms (10 loc + 10 scale,
nbases 20), 128 chains, one all-reduce(sum) of the [C, 1+P] logp+gradient block over NCCL.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
      tools/c5_obs_sharded.py [--n 50000000] [--steps 5] [--warmup 3]

Every rank generates ITS OWN slice of the observations (seeded per rank) and builds a handle
over that slice only; the knots, reparameterisations Z and penalties K come from one common
seeded subsample so that all ranks evaluate the same model.  Rank 0 adds the priors.  Timing:
CUDA events around (local sweep + all-reduce), barrier on both sides, max over ranks.
Checks: the reduced logp equals the sum of the per-rank partial logps gathered separately.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench_workload as bw
import liesel_ptm_b200 as lp
from liesel_ptm_b200 import parallel

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50_000_000)
ap.add_argument("--chains", type=int, default=128)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
args = ap.parse_args()

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n_loc = n_scale = 10
F32 = args.dtype == "f32"
DT = np.float32 if F32 else np.float64
T, nbases, C = n_loc + n_scale, 20, args.chains
lo, hi = parallel.obs_shard(args.n, world, rank)
n_local = hi - lo

# common subsample -> response map, knots, Z, K (identical on every rank)
y_sub, X_sub = bw.make_data(200_000, n_loc, n_scale, seed=1234)
bases = [lp.ps(X_sub[t].astype(DT), nbases=nbases, xname=f"x{t}", dtype=DT,
               knots=lp.equidistant_knots(np.array([-2.0, 2.0]), nbases, dtype=DT))
         for t in range(T)]

# this rank's slice (own seed); same generator as the other configs, response mapped by its
# own quantiles (synthetic data: only the shapes and value ranges matter)
y, X = bw.make_data(n_local, n_loc, n_scale, seed=100 + rank)
X = np.clip(X, -2.0 + 1e-9, 2.0 - 1e-9)
model = lp.LocScalePTM.new_ptm(y.astype(DT), nparam=20, to_float32=F32, device=f"cuda:{local}")
for t in range(T):
    b = lp.Basis(x=X[t].astype(DT), xname=f"x{t}", knots=bases[t].knots, Z=bases[t].Z, penalty=bases[t].penalty, nbases=nbases)
    tm = lp.term.f_ig(b, fname="f" if t < n_loc else "g")
    if t < n_loc:
        model.loc += tm
    else:
        model.scale += tm
model.build()
model.set_shard(0, n_local, add_priors=(rank == 0))
del X

# chain states: identical on every rank (seeded), small so that every term has sd ~ 0.1
sub_model = lp.LocScalePTM.new_ptm(y_sub.astype(DT), nparam=20, to_float32=F32, device=f"cuda:{local}")
for t in range(T):
    tm = lp.term.f_ig(bases[t], fname="f" if t < n_loc else "g")
    if t < n_loc:
        sub_model.loc += tm
    else:
        sub_model.scale += tm
params = torch.from_numpy(bw.make_states(sub_model, C, seed=7).astype(DT)).cuda()
P = params.shape[1]


def step():
    logp, grad = model.log_prob_and_grad(params)
    return parallel.allreduce_logp_grad(logp, grad)


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    logp, grad = step()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda", dtype=torch.float64)

# check: all-reduced logp == sum over ranks of the partials gathered one by one
part, _ = model.log_prob_and_grad(params)
part = part.clone()
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    gathered = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(gathered, part)
    ref = torch.stack(gathered).sum(0)
else:
    ref = part
err = float(((logp - ref).abs() / ref.abs().clamp_min(1.0)).max())
k_ms = None
model.profile(True)
model.log_prob_and_grad(params)
k_ms = model.last_kernel_ms()[1]
model.profile(False)
if rank == 0:
    out = {"config": "c5 observation-sharded", "n": args.n, "n_per_rank": n_local, "terms": T, "chains": C,
           "params_per_chain": P, "n_gpus": world, "ms_per_step": float(ms), "local_sweep_ms_rank0": k_ms,
           "evals_per_s": args.n * C / float(ms) * 1e3, "allreduce_bytes": C * (1 + P) * 8,
           "allreduce_check_rel_err": err, "finite": bool(torch.isfinite(logp).all() and torch.isfinite(grad).all())}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/r1_c5_n{world}.json", "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
