"""N-rank NCCL check of the observation-sharded wrappers: logp+grad, IWLS precision and
intercepts from shards + all-reduce equal the single-handle values on the full data.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/obs_shard_check.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from liesel_ptm_b200 import parallel
from tests.helpers import build_case, rel_err

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for dt in (np.float64, np.float32):
    full = build_case(n=20011, n_loc=2, n_scale=2, nbases=12, C=70, dtype=dt, device=f"cuda:{local}")
    params = torch.from_numpy(full.params).cuda()
    sh = parallel.ObsShardedLogProb(full.model)  # works on its own shard view of the model
    lp, g = sh.value_and_grad(params)
    lp0, g0 = full.model.log_prob_and_grad(params)
    K = torch.from_numpy(full.omodel.terms[0].K).cuda()
    P, L = sh.precision(0, params, K)
    P0, L0 = full.model.precision(0, params)
    b, gm = sh.intercepts(params, sequential=True)
    b0, gm0 = full.model.compute_intercepts(params, sequential=True)
    tag = np.dtype(dt).name
    out[tag] = {"logp": rel_err(lp.cpu().numpy(), lp0.cpu().numpy()), "grad": rel_err(g.cpu().numpy(), g0.cpu().numpy()),
                "precision": rel_err(P.cpu().numpy(), P0.cpu().numpy()), "chol": rel_err(L.cpu().numpy(), L0.cpu().numpy()),
                "beta0": float((b - b0).abs().max()), "gamma0": float((gm - gm0).abs().max()),
                "route": sh.model.last_xtwx_route(), "bounds": sh.bounds}
if rank == 0:
    print(json.dumps({"world": world, **out}))
    tol = {"float64": 1e-11, "float32": 5e-5}
    bad = [k for k, v in out.items() for m in ("logp", "grad", "precision", "chol") if not v[m] < tol[k]]
    sys.exit(1 if bad else 0)
if world > 1:
    dist.destroy_process_group()
