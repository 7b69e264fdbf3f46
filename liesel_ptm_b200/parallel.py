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
    """Sum per-rank partial logp [C] and grad [C, P] held in SEPARATE arrays with one
    all-reduce of a packed [C, 1 + P] block (two copy kernels; the product path,
    ``ObsShardedLogProb.value_and_grad``, has the epilogue kernel write the block directly)."""
    C = logp.shape[0]
    block = torch.empty((C, 1 + grad.shape[1]), dtype=grad.dtype, device=grad.device)
    block[:, 0] = logp
    block[:, 1:] = grad
    allreduce_block_(block, group)
    return block[:, 0], block[:, 1:]


def allreduce_block_(block: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the [C, 1 + P] block over the ranks of ``group``: one NCCL all-reduce
    on the current stream, nothing else (0.43 MB at config c5: latency-bound)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(block, op=dist.ReduceOp.SUM, group=group)
    return block


class ObsShardedLogProb:
    """logp+grad with the observations split across the ranks of ``group``.  Works on its own
    shard VIEW of the model (``LocScalePTM.shard_view``): the caller's model keeps evaluating
    all observations."""

    def __init__(self, model, group=None) -> None:
        self.full_model = model
        self.group = group
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        lo, hi = obs_shard(model.n, world, rank)
        self.model = model.shard_view(lo, hi, add_priors=(rank == 0))
        self.bounds = (lo, hi)
        self._block = None

    def value_and_grad(self, params: torch.Tensor):
        """(logp [C], grad [C, P]) over ALL observations: local sweep whose epilogue writes the
        packed block, one all-reduce in place, views of the block out."""
        C = params.shape[0]
        if self._block is None or self._block.shape[0] != C:
            self._block = torch.empty((C, 1 + self.model.num_params), dtype=self.model.dtype,
                                      device=self.model.device)
        block = self.model.log_prob_and_grad_block(params, self._block)
        allreduce_block_(block, self.group)
        return block[:, 0], block[:, 1:]

    def xtwx(self, term_index: int, params: torch.Tensor) -> torch.Tensor:
        """X'WX of a term over ALL observations: local sweep + one all-reduce of [C, p, p]."""
        x = self.model.xtwx(term_index, params)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(x, op=dist.ReduceOp.SUM, group=self.group)
        return x

    def precision(self, term_index: int, params: torch.Tensor, penalty: torch.Tensor):
        """IWLS precision and its Cholesky factor (iwls_proposals.py:82-89, 139-144) for
        observation-sharded data: the all-reduced X'WX, then penalty / ridge / factorisation
        (O(p^3) per chain) on every rank."""
        tau = params[:, self.model.layout["tau"][term_index]]
        return finish_precision(self.xtwx(term_index, params), penalty, tau)

    def intercepts(self, params: torch.Tensor, sequential: bool = False):
        """LocationIntercept / LogScaleIntercept.compute_intercept (predictor.py:80-100,
        126-130) over ALL observations: the five sums are additive over shards."""
        stats = allreduce_sum(self.model.intercept_stats(params), self.group)
        return intercepts_from_stats(stats, params, self.model.layout, self.model.n, sequential=sequential)


def allreduce_sum(x: torch.Tensor, group=None) -> torch.Tensor:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        x = x.clone()
        dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


def finish_precision(xtwx: torch.Tensor, penalty: torch.Tensor, tau: torch.Tensor):
    """precision = X'WX + K / max(tau, sqrt(eps))^2 + 1e-6 mean(diag) I and its lower
    Cholesky factor (iwls_proposals.py:82-89), batched over chains."""
    eps = torch.finfo(xtwx.dtype).eps ** 0.5
    inv_s2 = 1.0 / torch.clamp(tau, min=eps) ** 2
    P = xtwx + inv_s2[:, None, None] * penalty.to(xtwx)[None]
    p = P.shape[-1]
    ridge = 1e-6 * torch.diagonal(P, dim1=-2, dim2=-1).mean(-1)
    P = P + ridge[:, None, None] * torch.eye(p, dtype=P.dtype, device=P.device)[None]
    return P, torch.linalg.cholesky(P)


def intercepts_from_stats(stats: torch.Tensor, params: torch.Tensor, layout: dict, n_total: int,
                          sequential: bool = False):
    """(beta0_new, gamma0_new) from the (all-reduced) sums [C, 5] = (sum w, sum r, sum r^2,
    sum w^2, sum r w), w = exp(-eta_sigma), r = (y - eta_mu) w  (predictor.py:80-100, 126-130).
    ``sequential``: gamma0_new as LogScaleIntercept sees it when the location intercept has
    just been updated to beta0_new (its ``loc`` input is the full predictor,
    predictor.py:604-609): std of r - d w with d = beta0_new - beta0."""
    n = float(n_total)
    m_w, m_r, m_r2 = stats[:, 0] / n, stats[:, 1] / n, stats[:, 2] / n
    shift = m_r / m_w
    beta0 = params[:, layout["beta0"]] + shift
    if sequential:
        m_w2, m_rw = stats[:, 3] / n, stats[:, 4] / n
        var = m_r2 - 2.0 * shift * m_rw + shift * shift * m_w2 - (m_r - shift * m_w) ** 2
    else:
        var = m_r2 - m_r * m_r
    gamma0 = params[:, layout["gamma0"]] + 0.5 * torch.log(var)
    return beta0, gamma0
