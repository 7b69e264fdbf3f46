"""Multi-GPU plumbing for the hot path.  One process per GPU.

* chain axis (default): rank g owns chains [lo, hi); the static data is replicated;
  NO collective on the data path.
* observation axis (n >= 1e7): rank g owns observations [lo, hi) and evaluates every
  chain on its slice (``LocScalePTM.set_shard``); the partial ``[C, 1 + P]`` block
  (logp and gradient, priors added by rank 0 only) is summed with ONE all-reduce.

torch.distributed is the transport (NCCL over NVLink on the GPUs, gloo in the CPU
tests); the reduction payload for config c5 is 128 x 402 x 8 B = 0.41 MB, i.e. a
latency-bound single collective.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced split of ``total`` units: the first ``total % world`` ranks
    get one extra unit."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad world/rank: {world}/{rank}")
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def chain_shard(n_chains: int, world: int, rank: int) -> tuple[int, int]:
    return shard_bounds(n_chains, world, rank)


def obs_shard(n_obs: int, world: int, rank: int) -> tuple[int, int]:
    return shard_bounds(n_obs, world, rank)


def allreduce_logp_grad(logp: torch.Tensor, grad: torch.Tensor, group=None):
    """Sum the per-rank partial logp [C] and grad [C, P] with one all-reduce of the
    packed [C, 1 + P] block; returns (logp, grad) views of the reduced block."""
    C = logp.shape[0]
    block = torch.empty((C, 1 + grad.shape[1]), dtype=grad.dtype, device=grad.device)
    block[:, 0] = logp
    block[:, 1:] = grad
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(block, op=dist.ReduceOp.SUM, group=group)
    return block[:, 0], block[:, 1:]


class ObsShardedLogProb:
    """logp+grad with the observations split across the ranks of ``group``."""

    def __init__(self, model, group=None) -> None:
        self.model = model
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = obs_shard(model.n, world, rank)
        model.set_shard(lo, hi, add_priors=(rank == 0))
        self.bounds = (lo, hi)

    def value_and_grad(self, params: torch.Tensor):
        logp, grad = self.model.log_prob_and_grad(params)
        return allreduce_logp_grad(logp, grad, self.group)
