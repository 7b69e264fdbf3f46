"""Observation-axis sharding over NCCL on real GPUs: skipped below 2 GPUs.

Launches tools/obs_shard_check.py under torch.distributed.run: every rank evaluates its shard
view for all chains; the packed [C, 1 + P] block written by the epilogue kernel is summed
with ONE NCCL all-reduce; logp + gradient, IWLS precision + Cholesky (all-reduced X'WX) and
the intercept statistics must equal the single-handle values over all observations.
"""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_obs_sharded_nccl_two_ranks():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + (os.getpid() % 300)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tools", "obs_shard_check.py")]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2
    assert out["float64"]["logp"] < 1e-12 and out["float64"]["grad"] < 1e-11
    assert out["float32"]["logp"] < 1e-5
