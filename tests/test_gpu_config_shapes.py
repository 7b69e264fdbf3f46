"""Oracle parity at the shapes BASELINE.json's configs name.

The CUDA path runs the configuration itself (c2 in full; c3 with its 1024 chains at n = 1M;
c4 and the c5 per-GPU shape with their term counts and basis sizes); the NumPy oracle --
dense designs, 7.6e6 evals/s on 16 cores -- is evaluated on a subset of the chain batch
(chains are independent, so a subset pins every chain's arithmetic) at the FULL number of
observations.  Tolerances are the north-star ones: rel 1e-12 float64, rel 1e-5 float32 on the
f32-rounded inputs, with the parity definition of SURVEY.md section 7 (|dlogp| / |logp| and
|dg|_inf / |g|_inf per chain).
"""

from __future__ import annotations

import functools

import numpy as np
import pytest
import torch

import bench_workload as bw
import liesel_ptm_b200 as lp
from tests.helpers import build_case, oracle_eval, oracle_from_workload, pack_grad_layout, rel_err

pytestmark = pytest.mark.gpu

C3 = "c3_n1M_5loc_5scale_nb20_1024chains"
C2 = "c2_n100k_3loc_1scale_64chains"
C4 = "c4_n1M_10loc_1scale_nb30_256chains"


def _per_chain_err(a, b):
    """max over chains of |a_c - b_c|_inf / |b_c|_inf."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
    ax = tuple(range(1, a.ndim))
    return float(np.max(np.abs(a - b).max(axis=ax) / np.maximum(np.abs(b).max(axis=ax), 1e-300)))


def _oracle_lp_grad(om, st, L, P):
    lp_ref, g = om.logp_and_grad(st)
    return lp_ref, pack_grad_layout(L, P, g)


def _f32_model_from(wl64):
    """float32 problem holding the f32-rounded data of a float64 workload (same Z, K, knots)."""
    m = lp.LocScalePTM(wl64.y.astype(np.float32), wl64.model.knots.astype(np.float32), to_float32=True)
    for t, k in wl64.model.all_terms():
        b = t.basis
        tm = lp.term.f_ig(lp.Basis(x=b.x.astype(np.float32), xname=b.xname, knots=b.knots.astype(np.float32),
                                   Z=np.ascontiguousarray(b.Z.astype(np.float32)),
                                   penalty=np.ascontiguousarray(b.penalty.astype(np.float32)),
                                   nbases=b.nbases), fname="f" if k == 0 else "g")
        tm.penalty_rank = t.penalty_rank
        if k == 0:
            m.loc += tm
        else:
            m.scale += tm
    if wl64.model.trafo_Z is not None:
        m.trafo_Z = wl64.model.trafo_Z.astype(np.float32)
    return m.build()


# ---------------------------------------------------------------------------------------
# c2: n = 100k, 3 loc + 1 scale, 64 chains -- the whole configuration against the oracle
# ---------------------------------------------------------------------------------------
def test_c2_full_config_f64():
    wl = bw.build(C2, dtype=np.float64)
    model = wl.model.build()
    assert wl.n == 100_000 and wl.params.shape[0] == 64
    params = torch.from_numpy(wl.params).cuda()
    logp, grad = model.log_prob_and_grad(params)
    om, st, L, P = oracle_from_workload(wl)
    lp_ref, g_ref = _oracle_lp_grad(om, st, L, P)
    assert _per_chain_err(logp.cpu().numpy(), lp_ref) < 1e-12
    assert _per_chain_err(grad.cpu().numpy(), g_ref) < 1e-12
    assert _per_chain_err(model.log_prob(params).cpu().numpy(), lp_ref) < 1e-12


def test_c2_full_config_f32():
    """float32 problem, all 64 chains on the GPU, the f64 oracle on the f32-rounded numbers
    for 16 of them: rel 1e-5 on logp and on the gradient (inf-norm per chain)."""
    wl = bw.build(C2, dtype=np.float32)
    model = wl.model.build()
    params = torch.from_numpy(wl.params).cuda()
    logp, grad = model.log_prob_and_grad(params)
    idx = np.arange(0, 64, 4)
    om, st, L, P = oracle_from_workload(wl, idx, as_f64=True)
    lp_ref, g_ref = _oracle_lp_grad(om, st, L, P)
    assert _per_chain_err(logp.cpu().numpy()[idx], lp_ref) < 1e-5
    assert _per_chain_err(grad.cpu().numpy()[idx], g_ref) < 1e-5


# ---------------------------------------------------------------------------------------
# c3: n = 1M, 5 + 5 terms, 1024 chains (the bench configuration); oracle on 4 of the chains
# ---------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=1)
def _c3():
    wl = bw.build(C3, dtype=np.float64)
    model = wl.model.build()
    idx = np.array([0, 341, 682, 1023])
    om, st, L, P = oracle_from_workload(wl, idx)
    return wl, model, idx, om, st, L, P


def test_c3_logp_grad_vs_oracle():
    wl, model, idx, om, st, L, P = _c3()
    assert wl.n == 1_000_000 and wl.params.shape[0] == 1024
    params = torch.from_numpy(wl.params).cuda()
    logp, grad = model.log_prob_and_grad(params)
    lp_ref, g_ref = _oracle_lp_grad(om, st, L, P)
    assert _per_chain_err(logp.cpu().numpy()[idx], lp_ref) < 1e-12
    assert _per_chain_err(grad.cpu().numpy()[idx], g_ref) < 1e-12
    assert _per_chain_err(model.log_prob(params).cpu().numpy()[idx], lp_ref) < 1e-12
    # the packed [C, 1 + P] block of the observation-axis all-reduce carries the same bits
    block = model.log_prob_and_grad_block(params)
    assert torch.equal(block[:, 0], logp) and torch.equal(block[:, 1:], grad)


def test_c3_information_intercepts_pointwise_vs_oracle():
    wl, model, idx, om, st, L, P = _c3()
    sub = torch.from_numpy(wl.params[idx]).cuda()
    # IWLS information of all five location terms from one sweep (iwls_proposals.py:55-89)
    Ps, Ls = model.precision_all(sub)
    for t in range(wl.n_loc):
        _, P_ref, L_ref = om.loc_precision(st, t)
        assert _per_chain_err(Ps[t].cpu().numpy(), P_ref) < 1e-11
        assert _per_chain_err(Ls[t].cpu().numpy(), L_ref) < 1e-9
    # intercept pseudo-parameters, both evaluation orders (predictor.py:80-100, 126-130, 604-609)
    for seq in (False, True):
        b0, g0 = model.compute_intercepts(sub, sequential=seq)
        ref = np.array([om.compute_intercepts(st, c, sequential=seq) for c in range(len(idx))])
        assert np.max(np.abs(b0.cpu().numpy() - ref[:, 0])) < 1e-11
        assert np.max(np.abs(g0.cpu().numpy() - ref[:, 1])) < 1e-11
    # pointwise records over the four samples (evaluate.py:106-133)
    ll = np.stack([om.loglik_terms(st, c)["ll"] for c in range(len(idx))])
    mx = ll.max(axis=0)
    lppd_ref = mx + np.log(np.exp(ll - mx).sum(axis=0)) - np.log(len(idx))
    lppd, pwaic = model.pointwise(sub)
    assert rel_err(lppd.cpu().numpy(), lppd_ref) < 1e-12
    assert np.allclose(pwaic.cpu().numpy(), ll.var(axis=0), rtol=1e-8, atol=1e-12)


# ---------------------------------------------------------------------------------------
# c4: n = 1M, 10 location terms with 30 bases + 1 scale term, 256 chains; oracle on 4 chains
# ---------------------------------------------------------------------------------------
def test_c4_shape_logp_grad_and_information():
    wl = bw.build(C4, dtype=np.float64)
    model = wl.model.build()
    assert wl.n == 1_000_000 and wl.params.shape[0] == 256 and wl.nbases == 30
    idx = np.array([0, 85, 170, 255])
    om, st, L, P = oracle_from_workload(wl, idx)
    params = torch.from_numpy(wl.params).cuda()
    logp, grad = model.log_prob_and_grad(params)
    lp_ref, g_ref = _oracle_lp_grad(om, st, L, P)
    assert _per_chain_err(logp.cpu().numpy()[idx], lp_ref) < 1e-12
    assert _per_chain_err(grad.cpu().numpy()[idx], g_ref) < 1e-12
    # X'WX of all ten location terms: float64 banded sweep ...
    sub = torch.from_numpy(wl.params[idx]).cuda()
    refs = [om.loc_precision(st, t) for t in range(wl.n_loc)]
    Ps, Ls = model.precision_all(sub)
    assert model.last_xtwx_route() == "banded"
    for t in range(wl.n_loc):
        assert _per_chain_err(Ps[t].cpu().numpy(), refs[t][1]) < 1e-11
        assert _per_chain_err(Ls[t].cpu().numpy(), refs[t][2]) < 1e-9
    x0 = model.xtwx(3, sub)
    assert _per_chain_err(x0.cpu().numpy(), refs[3][0]) < 1e-12
    # ... and the float32 tcgen05 route (3xTF32) on the f32-rounded data of the same workload
    m32 = _f32_model_from(wl)
    sub32 = torch.from_numpy(wl.params[idx].astype(np.float32)).cuda()
    Ps32, _ = m32.precision_all(sub32)
    assert m32.last_xtwx_route().startswith("tcgen05")
    for t in range(wl.n_loc):
        assert _per_chain_err(Ps32[t].cpu().numpy(), refs[t][1]) < 2e-5
    x32 = m32.xtwx(3, sub32)
    assert _per_chain_err(x32.cpu().numpy(), refs[3][0]) < 2e-5


# ---------------------------------------------------------------------------------------
# c5 per-GPU shape: 10 + 10 terms (two terms per term warp), 128 chains, n = 200k
# ---------------------------------------------------------------------------------------
def test_c5_shape_logp_grad_two_terms_per_warp():
    bw.CONFIGS["_c5_shape_n200k"] = (200_000, 10, 10, 20, 128)
    wl = bw.build("_c5_shape_n200k", dtype=np.float64)
    model = wl.model.build()
    params = torch.from_numpy(wl.params).cuda()
    logp, grad = model.log_prob_and_grad(params)
    idx = np.arange(0, 128, 16)
    om, st, L, P = oracle_from_workload(wl, idx)
    lp_ref, g_ref = _oracle_lp_grad(om, st, L, P)
    assert _per_chain_err(logp.cpu().numpy()[idx], lp_ref) < 1e-12
    assert _per_chain_err(grad.cpu().numpy()[idx], g_ref) < 1e-12


# ---------------------------------------------------------------------------------------
# stress shapes (formerly tools/big_check.py)
# ---------------------------------------------------------------------------------------
def test_5000_chains_vs_oracle():
    """157 chain groups, ragged last group: every 25th chain against the oracle."""
    case = build_case(n=1500, n_loc=2, n_scale=1, nbases=12, C=5000)
    p = torch.from_numpy(case.params).cuda()
    lpv, g = case.model.log_prob_and_grad(p)
    lp_ref, g_ref = oracle_eval(case)
    assert _per_chain_err(lpv.cpu().numpy(), lp_ref) < 1e-12
    assert _per_chain_err(g.cpu().numpy(), g_ref) < 1e-12
    Pm, _ = case.model.precision(0, p)
    _, rP, _ = case.omodel.loc_precision(case.state, 0)
    assert _per_chain_err(Pm.cpu().numpy(), rP) < 1e-11


def test_30M_observations_shards_add_up():
    """n = 3e7 > 2^24 observations (64-bit observation indexing): ragged shards, including
    single-observation ones at both ends, add up to the full evaluation."""
    n = 30_000_000
    y, X = bw.make_data(n, 1, 1, seed=3)
    m = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=False)
    kn = lp.equidistant_knots(np.array([-2.0, 2.0]), 20)
    sub = slice(0, 200_000)
    for t in range(2):
        b0 = lp.ps(X[t][sub], nbases=20, xname=f"x{t}", knots=kn)
        b = lp.Basis(x=np.clip(X[t], -2 + 1e-9, 2 - 1e-9), xname=f"x{t}", knots=b0.knots, Z=b0.Z,
                     penalty=b0.penalty, nbases=20)
        (m.loc if t == 0 else m.scale).__iadd__(lp.term.f_ig(b, fname="f" if t == 0 else "g"))
    m.build()
    par = torch.from_numpy(bw.make_states(m, 40, seed=2)).cuda()
    full, gfull = m.log_prob_and_grad(par)
    full, gfull = full.clone(), gfull.clone()
    acc, gacc = torch.zeros_like(full), torch.zeros_like(gfull)
    for lo, hi, pri in ((0, 1, True), (1, 16_777_217, False), (16_777_217, n - 1, False), (n - 1, n, False)):
        v = m.shard_view(lo, hi, add_priors=pri)
        a, b_ = v.log_prob_and_grad(par)
        acc += a
        gacc += b_
    assert float(((acc - full).abs() / full.abs()).max()) < 1e-12
    assert float((gacc - gfull).abs().max() / gfull.abs().max()) < 1e-11
    # a view is a new handle: the model it came from still evaluates everything
    again, _ = m.log_prob_and_grad(par)
    assert torch.equal(again, full)
    # shard-local precision / Cholesky are refused (not additive over shards)
    v = m.shard_view(0, n // 2, add_priors=False)
    with pytest.raises(lp.PtmError):
        v.precision(0, par)
    assert v.xtwx(0, par).shape == (40, 19, 19)
    with pytest.raises(RuntimeError):
        v.compute_intercepts(par)
