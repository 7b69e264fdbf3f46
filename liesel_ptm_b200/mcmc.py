"""Batched MCMC transitions over the fused operator (host glue, torch on the GPU).

This is the *caller* side of the hot path, restated so that the reference's headline
"ESS/s" can be measured without liesel.goose: every transition below touches the model
only through ``log_prob_and_grad`` / ``precision`` (the C ABI), for all chains at once.

* ``iwls_transition`` follows ``IWLSKernelFixed._standard_transition``
  (iwls_fixed.py:145-188) and ``gs.IWLSKernel`` (same algorithm with
  ``chol_info_fn(state)`` evaluated at the current and the proposed state,
  iwls_proposals.py:82-89): proposal mean ``pos + s^2/2 * P^-1 score``, a Gaussian
  draw with precision-Cholesky ``chol / s``, the reverse density, Metropolis-Hastings.
  ``solve`` / ``mvn_sample`` / ``mvn_log_prob`` restate liesel.goose.iwls_utils (not in
  tree): precision-Cholesky parameterisation, log-density without the 2*pi constant
  cancelling in the ratio.
* ``hmc_transition`` / ``mala_transition`` are plain HMC / Langevin steps for the block
  without an information matrix (the transformation coefficients delta_z; the reference
  samples them with NUTS from blackjax, which is out of scope).

* ``gibbs_tau`` is the inverse-gamma Gibbs update of the smoothing variances
  (``star_ig_gibbs``, gam/kernel.py:9-41: ``tau^2 ~ b' / Gamma(a')``, ``a' = a + rank / 2``,
  ``b' = b + beta' K beta / 2``) over the fused quadratic forms (``ptm_quadform``), and
  ``update_intercepts`` the pseudo-Gibbs recomputation of the two intercepts
  (``loc_intercept_pgibbs`` / ``log_scale_intercept_pgibbs``, predictor.py:133-175, 239-281)
  from the fused sufficient statistics in the sampler's sequential order.  With both switched
  on ``run_chains`` runs the composition of the reference's default strategy except for the
  blocks that have no information matrix (delta_z: HMC here, NUTS there; tau_delta: fixed).
"""

from __future__ import annotations

import dataclasses
import time

import numpy as np
import torch

from . import _lib
from .model import LocScalePTM


def _solve(chol: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """P^-1 rhs from the lower Cholesky factor of P (iwls_utils.solve)."""
    # two triangular solves (cuBLAS trsm, capturable in a CUDA graph; cholesky_solve goes
    # through MAGMA, which allocates)
    y = torch.linalg.solve_triangular(chol, rhs.unsqueeze(-1), upper=False)
    return torch.linalg.solve_triangular(chol.transpose(-1, -2), y, upper=True).squeeze(-1)


def _mvn_sample(gen, mean: torch.Tensor, chol_prec: torch.Tensor) -> torch.Tensor:
    """mean + L^-T z for precision P = L L' (iwls_utils.mvn_sample)."""
    z = torch.randn(mean.shape, dtype=mean.dtype, device=mean.device, generator=gen)
    return mean + torch.linalg.solve_triangular(chol_prec.transpose(-1, -2), z.unsqueeze(-1), upper=True).squeeze(-1)


def _mvn_log_prob(x: torch.Tensor, mean: torch.Tensor, chol_prec: torch.Tensor) -> torch.Tensor:
    """log N(x; mean, (L L')^-1) up to the 2*pi constant (iwls_utils.mvn_log_prob)."""
    zq = ((x - mean).unsqueeze(-2) @ chol_prec).squeeze(-2)
    return -0.5 * (zq * zq).sum(-1) + torch.log(torch.diagonal(chol_prec, dim1=-2, dim2=-1)).sum(-1)


@dataclasses.dataclass
class TransitionInfo:
    accepted: torch.Tensor  # [C] bool
    log_alpha: torch.Tensor  # [C]


def _mh(gen, params, prop, logp, logp_prop, correction):
    log_alpha = logp_prop - logp + correction
    log_alpha = torch.where(torch.isnan(log_alpha), torch.full_like(log_alpha, -float("inf")), log_alpha)
    u = torch.rand(logp.shape, dtype=logp.dtype, device=logp.device, generator=gen)
    acc = torch.log(u) < log_alpha
    return torch.where(acc[:, None], prop, params), acc, log_alpha


def iwls_transition(model: LocScalePTM, params: torch.Tensor, term_index: int, step_size: float, gen,
                    cache=None, chol=None):
    """One IWLS-MH transition of the coefficient block of ``term_index`` for all chains.
    Cost per call: 2 logp+grad sweeps and 2 information sweeps (iwls_fixed.py:159-186).

    ``chol``: precision Cholesky factor [C, p, p] valid at the current AND the proposed
    state.  The information of a location term depends on the state only through the scale
    predictor (w = 1/sigma^2) and tau, that of a scale term (W == 2) only on tau, so a
    proposal that moves this block alone leaves it unchanged; passing it skips both
    information sweeps without changing the transition."""
    lo = model.layout["beta"][term_index]
    p = model.terms[term_index][0].basis.ncoef
    s = step_size
    if torch.is_tensor(s):  # per-chain step sizes
        s = s[:, None]
    if cache is None:
        logp, grad = model.log_prob_and_grad(params)
        logp, grad = logp.clone(), grad.clone()
    else:
        logp, grad = cache
    fixed_info = chol is not None
    if not fixed_info:
        _, chol = model.precision(term_index, params)
    pos = params[:, lo: lo + p]
    mu = pos + 0.5 * s * s * _solve(chol, grad[:, lo: lo + p])
    s3 = s[:, :, None] if torch.is_tensor(s) else s
    prop_block = _mvn_sample(gen, mu, chol / s3)
    fwd = _mvn_log_prob(prop_block, mu, chol / s3)
    prop = params.clone()
    prop[:, lo: lo + p] = prop_block
    logp_p, grad_p = model.log_prob_and_grad(prop)
    logp_p, grad_p = logp_p.clone(), grad_p.clone()
    chol_p = chol if fixed_info else model.precision(term_index, prop)[1]
    mu_p = prop_block + 0.5 * s * s * _solve(chol_p, grad_p[:, lo: lo + p])
    bwd = _mvn_log_prob(pos, mu_p, chol_p / s3)
    new, acc, la = _mh(gen, params, prop, logp, logp_p, bwd - fwd)
    new_logp = torch.where(acc, logp_p, logp)
    new_grad = torch.where(acc[:, None], grad_p, grad)
    return new, (new_logp, new_grad), TransitionInfo(acc, la)


def mala_transition(model: LocScalePTM, params: torch.Tensor, lo: int, width: int, step_size, gen,
                    cache=None):
    """Metropolis-adjusted Langevin step on params[:, lo:lo+width] (identity metric)."""
    if cache is None:
        logp, grad = model.log_prob_and_grad(params)
        logp, grad = logp.clone(), grad.clone()
    else:
        logp, grad = cache
    s = step_size if torch.is_tensor(step_size) else torch.full((params.shape[0],), float(step_size),
                                                                 dtype=params.dtype, device=params.device)
    s = s[:, None]
    pos = params[:, lo: lo + width]
    mu = pos + 0.5 * s * s * grad[:, lo: lo + width]
    z = torch.randn(pos.shape, dtype=pos.dtype, device=pos.device, generator=gen)
    prop_block = mu + s * z
    prop = params.clone()
    prop[:, lo: lo + width] = prop_block
    logp_p, grad_p = model.log_prob_and_grad(prop)
    logp_p, grad_p = logp_p.clone(), grad_p.clone()
    mu_p = prop_block + 0.5 * s * s * grad_p[:, lo: lo + width]
    fwd = -0.5 * (((prop_block - mu) / s) ** 2).sum(-1)
    bwd = -0.5 * (((pos - mu_p) / s) ** 2).sum(-1)
    new, acc, la = _mh(gen, params, prop, logp, logp_p, bwd - fwd)
    return new, (torch.where(acc, logp_p, logp), torch.where(acc[:, None], grad_p, grad)), TransitionInfo(acc, la)


def effective_sample_size(x: np.ndarray) -> float:
    """Bulk ESS of draws x [chains, samples] (Geyer's initial positive sequence on the
    chain-averaged autocovariance, as arviz / Stan do; rank normalisation omitted)."""
    x = np.asarray(x, dtype=np.float64)
    m, n = x.shape
    xc = x - x.mean(axis=1, keepdims=True)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(xc, nfft, axis=1)
    acov = np.fft.irfft(f * np.conj(f), nfft, axis=1)[:, :n] / n
    w = acov[:, 0].mean() * n / (n - 1)
    b = x.mean(axis=1).var(ddof=1) if m > 1 else 0.0
    var_plus = w * (n - 1) / n + b
    if var_plus <= 0:
        return float(m * n)
    rho = 1.0 - (w - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    tau, t = -1.0, 0
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        tau += 2.0 * pair
        t += 2
    return float(m * n / max(tau, 1.0 / np.log10(max(m * n, 10))))


def hmc_transition(model: LocScalePTM, params: torch.Tensor, lo: int, width: int, step_size, n_leapfrog: int,
                   metric_chol: torch.Tensor, gen, cache=None):
    """Hamiltonian Monte Carlo on params[:, lo:lo+width] with a dense metric: ``metric_chol``
    is the lower Cholesky factor L of the inverse mass matrix Sigma = L L' (the pooled
    posterior covariance estimate), per-chain step sizes, ``n_leapfrog`` logp+grad sweeps.
    Stands in for the reference's NUTS block on delta_z (blackjax, out of scope)."""
    if cache is None:
        logp, grad = model.log_prob_and_grad(params)
        logp, grad = logp.clone(), grad.clone()
    else:
        logp, grad = cache
    eps = (step_size if torch.is_tensor(step_size) else
           torch.full((params.shape[0],), float(step_size), dtype=params.dtype, device=params.device))[:, None]
    Lm = metric_chol
    sigma = Lm @ Lm.T
    z = torch.randn((params.shape[0], width), dtype=params.dtype, device=params.device, generator=gen)
    mom0 = torch.linalg.solve_triangular(Lm.T, z.T, upper=True).T  # p = L^-T z ~ N(0, Sigma^-1)
    mom = mom0 + 0.5 * eps * grad[:, lo: lo + width]
    q = params.clone()
    logp_q, grad_q = logp, grad
    for i in range(n_leapfrog):
        q[:, lo: lo + width] += eps * (mom @ sigma)
        logp_q, grad_q = model.log_prob_and_grad(q)
        logp_q, grad_q = logp_q.clone(), grad_q.clone()
        mom = mom + (0.5 if i == n_leapfrog - 1 else 1.0) * eps * grad_q[:, lo: lo + width]
    kin0 = 0.5 * (z * z).sum(-1)
    kin1 = 0.5 * ((mom @ Lm) ** 2).sum(-1)
    new, acc, la = _mh(gen, params, q, logp, logp_q, kin0 - kin1)
    return new, (torch.where(acc, logp_q, logp), torch.where(acc[:, None], grad_q, grad)), TransitionInfo(acc, la)


def gibbs_tau(model: LocScalePTM, params: torch.Tensor, gen) -> torch.Tensor:
    """``star_ig_gibbs`` (gam/kernel.py:9-41) for every smooth term and chain at once:
    ``tau_t^2 = (b + beta' K beta / 2) / Gamma(a + rank / 2)`` with the term's inverse-gamma
    hyper-parameters (gam/var.py:293-294, 352-360).  Non-centred terms are left alone (their
    scale is sampled jointly with the latent coefficients in the reference).  Returns a new
    parameter block."""
    T = len(model.terms)
    if T == 0:
        return params
    L = model.layout
    quad = model.quadform(params)  # [C, T]
    a = torch.tensor([t.ig_concentration + 0.5 * t.penalty_rank for t, _ in model.terms], dtype=params.dtype,
                     device=params.device)
    b = torch.tensor([t.ig_scale for t, _ in model.terms], dtype=params.dtype, device=params.device)
    g = torch._standard_gamma(a.expand(params.shape[0], T).contiguous(), generator=gen)
    tau = torch.sqrt((b + 0.5 * quad) / g)
    out = params.clone()
    keep = torch.tensor([bool(t.is_noncentered) for t, _ in model.terms], device=params.device)
    cols = slice(L["tau"][0], L["tau"][0] + T)
    out[:, cols] = torch.where(keep, params[:, cols], tau)
    return out


def update_intercepts(model: LocScalePTM, params: torch.Tensor) -> torch.Tensor:
    """``loc_intercept_pgibbs`` then ``log_scale_intercept_pgibbs`` (predictor.py:133-175,
    239-281): both intercepts recomputed from the current smooth terms, the log-scale
    intercept seeing the updated location intercept (one fused sweep, ``ptm_intercept_stats``)."""
    b0, g0 = model.compute_intercepts(params, sequential=True)
    out = params.clone()
    out[:, model.layout["beta0"]] = b0
    out[:, model.layout["gamma0"]] = g0
    return out


class _NoShapeBlock:
    """Stand-in for the HMC block's record when the model has no shape parameters."""

    def __init__(self, params):
        self.log_alpha = torch.zeros(params.shape[0], dtype=params.dtype, device=params.device)
        self.accepted = torch.ones(params.shape[0], dtype=torch.bool, device=params.device)


def run_chains(model: LocScalePTM, params0: torch.Tensor, warmup: int, draws: int, seed: int = 0,
               iwls_step: float = 1.0, dz_step: float = 0.05, target_accept: float = 0.8, n_leapfrog: int = 8,
               reuse_information: bool = True, use_graph: bool = True, sample_tau: bool = False,
               pgibbs_intercepts: bool = False):
    """Sweep per iteration: an IWLS-MH transition per smooth term, an HMC transition on
    delta_z.  Warm-up adapts the per-chain step sizes (Robbins-Monro on the acceptance
    probability) and, at 2/8 ... 7/8 of its length, the dense metric of delta_z from the
    pooled across-chain covariance of the draws since the previous update.  With ``reuse_information`` the information matrices are
    evaluated once per sweep for all location terms (``precision_all``; they only change
    when a scale term or an intercept moves) and once per run for the scale terms (W == 2,
    tau fixed) instead of twice per transition -- same transitions, fewer sweeps.  With
    ``use_graph`` the posterior epoch replays one captured CUDA graph per iteration.
    Returns (draws [draws, C, P] on the host, info dict with timings and acceptance)."""
    gen = torch.Generator(device=params0.device)
    gen.manual_seed(seed)
    params = params0.clone()
    C = params.shape[0]
    L = model.layout
    dz_lo, dz_w = L["delta_z"], model.nparam - 1
    step = torch.full((C,), float(dz_step), dtype=params.dtype, device=params.device)
    metric_chol = torch.eye(dz_w, dtype=params.dtype, device=params.device)
    adapt_t = 0
    has_shape = getattr(model, "bspline", "ptm") != "identity"
    metric_updates = {(warmup * k) // 8 for k in (2, 3, 4, 5, 6, 7)} if has_shape else set()
    T = len(model.terms)
    is_loc = [pred == _lib.PTM_LOC for _, pred in model.terms]
    istep = torch.full((max(T, 1), C), float(iwls_step), dtype=params.dtype, device=params.device)
    cache = None
    out = torch.empty((draws, C, params.shape[1]), dtype=params.dtype, device=params.device)
    scale_chol = {}
    scale_xtx, scale_pen = {}, {}
    if reuse_information:  # W == 2: X'WX of a scale term is constant, its factor only moves with tau
        for t in range(T):
            if not is_loc[t]:
                if sample_tau:
                    scale_xtx[t] = model.xtwx(t, params).clone()
                    scale_pen[t] = torch.from_numpy(np.asarray(model.terms[t][0].penalty)).to(params)
                else:
                    scale_chol[t] = model.precision(t, params)[1].clone()
    window = []

    def sweep(params, cache, it_warm, gen=gen):
        """One iteration at the current step sizes / metric; returns the new state and the
        acceptance indicators of the IWLS blocks (mean over blocks) and of the HMC block."""
        nonlocal istep, step, adapt_t
        loc_chol = None
        if reuse_information and any(is_loc):
            loc_chol = [c.clone() for c in model.precision_all(params)[1]]
        k_loc = 0
        acc_i = torch.zeros((), dtype=torch.float64, device=params.device)
        for t in range(T):
            chol = None
            if reuse_information:
                if is_loc[t]:
                    chol = loc_chol[k_loc]
                elif t in scale_chol:
                    chol = scale_chol[t]
                else:  # tau moves: penalty, ridge and factor on the constant 2 X'X
                    from .parallel import finish_precision

                    chol = finish_precision(scale_xtx[t], scale_pen[t], params[:, L["tau"][t]])[1]
                    if model.terms[t][0].is_noncentered:
                        chol = chol * params[:, L["tau"][t]][:, None, None]
            k_loc += 1 if is_loc[t] else 0
            params, cache, info = iwls_transition(model, params, t, istep[t], gen, cache, chol=chol)
            if it_warm is not None:
                a = torch.exp(torch.clamp(info.log_alpha, max=0.0))
                istep[t] = torch.clamp(istep[t] * torch.exp((a - 0.8) * (10.0 / (it_warm + 10.0))), 1e-3, 2.0)
            else:
                acc_i = acc_i + info.accepted.double().mean() / T
        if has_shape:
            params, cache, info = hmc_transition(model, params, dz_lo, dz_w, step, n_leapfrog, metric_chol, gen, cache)
        else:  # Gaussian model (bspline="identity", model/model.py:386-417): no shape parameters to move
            info = _NoShapeBlock(params)
        if sample_tau or pgibbs_intercepts:
            # Gibbs blocks of the reference's composition; the cached logp / gradient are stale after
            # them and are refreshed with one fused sweep
            if sample_tau:
                params = gibbs_tau(model, params, gen)
            if pgibbs_intercepts:
                params = update_intercepts(model, params)
            lp_n, g_n = model.log_prob_and_grad(params)
            cache = (lp_n.clone(), g_n.clone())
        return params, cache, info, acc_i

    # ---- warm-up (eager: host-side adaptation logic) -----------------------------------
    for it in range(warmup):
        params, cache, info, _ = sweep(params, cache, it)
        a = torch.exp(torch.clamp(info.log_alpha, max=0.0))
        step = step * torch.exp((a - target_accept) * (10.0 / (adapt_t + 10.0)))
        adapt_t += 1
        if has_shape and it >= warmup // 8:
            window.append(params[:, dz_lo: dz_lo + dz_w].clone())
        if window and (it + 1) in metric_updates and C * len(window) >= 10 * dz_w:
            cov = torch.cov(torch.cat(window, 0).T)
            cov = cov + 1e-6 * torch.diagonal(cov).mean() * torch.eye(dz_w, dtype=cov.dtype, device=cov.device)
            metric_chol = torch.linalg.cholesky(cov)
            step = torch.full_like(step, 0.5 * dz_w ** -0.25)  # whitened scale; re-adapt from there
            adapt_t = 0
            window = []
    if cache is None:
        lp0, g0 = model.log_prob_and_grad(params)
        cache = (lp0.clone(), g0.clone())

    # ---- posterior epoch: the whole sweep is one CUDA graph (fixed step sizes and metric;
    #      the library enqueues on the capturing stream, no host synchronisation inside) ----
    acc_i_sum = torch.zeros((), dtype=torch.float64, device=params.device)
    acc_d_sum = torch.zeros((), dtype=torch.float64, device=params.device)
    graph, graph_error = None, None
    if use_graph and draws > 0:
        try:
            g_params, g_logp, g_grad = params.clone(), cache[0].clone(), cache[1].clone()
            # the graph owns its generator: a generator registered with a graph is only
            # usable inside captures afterwards
            gen_g = torch.Generator(device=params0.device)
            gen_g.manual_seed(seed + 1)
            side = torch.cuda.Stream(device=params.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                sweep(g_params, (g_logp, g_grad), None)  # allocate workspaces outside the capture
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            graph.register_generator_state(gen_g)
            with torch.cuda.graph(graph, stream=side):
                p1, c1, info1, ai = sweep(g_params, (g_logp, g_grad), None, gen_g)
                g_params.copy_(p1)
                g_logp.copy_(c1[0])
                g_grad.copy_(c1[1])
                acc_i_sum += ai
                acc_d_sum += info1.accepted.double().mean()
            acc_i_sum.zero_()
            acc_d_sum.zero_()
            g_params.copy_(params)
            g_logp.copy_(cache[0])
            g_grad.copy_(cache[1])
        except Exception as exc:  # capture not possible (old torch, unsupported op): eager sweeps
            graph = None
            graph_error = repr(exc)
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(draws):
        if graph is not None:
            graph.replay()
            out[it].copy_(g_params)
        else:
            params, cache, info, ai = sweep(params, cache, None)
            acc_i_sum += ai
            acc_d_sum += info.accepted.double().mean()
            out[it] = params
    torch.cuda.synchronize()
    t_post = time.perf_counter() - t0
    n_it = max(draws, 1)
    return out.cpu().numpy(), {"posterior_seconds": t_post, "accept_iwls": float(acc_i_sum) / n_it,
                               "accept_dz": float(acc_d_sum) / n_it, "dz_step_median": float(step.median()),
                               "n_leapfrog": n_leapfrog, "cuda_graph": graph is not None,
                               "cuda_graph_error": graph_error}
