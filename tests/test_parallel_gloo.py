"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain sharding needs no
collective; observation sharding sums one packed [C, 1+P] block."""

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from liesel_ptm_b200.parallel import (  # noqa: E402
    allreduce_logp_grad, allreduce_sum, chain_shard, finish_precision, intercepts_from_stats, obs_shard,
    shard_bounds,
)


def test_shard_bounds_cover_everything():
    for total in (0, 1, 7, 1024, 1_000_003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(total, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ptm_oracle as o
    from tests.helpers import build_case, pack_grad

    case = build_case(n=240, n_loc=1, n_scale=1, nbases=10, C=4, build=False)
    om, st = case.omodel, case.state
    lp_full, g_full = om.logp_and_grad(st)
    g_full = pack_grad(case, g_full)

    # --- observation-axis sharding: partial sums + one all-reduce ---------------------
    lo, hi = obs_shard(om.n, world, rank)
    terms = [o.term_from_reparam(t.x[lo:hi], t.knots, t.Z, t.K, t.rank, t.predictor) for t in om.terms]
    sub = o.PTMModel(om.y[lo:hi], terms, nparam=om.nparam)
    sub.Zd = om.Zd
    lp_s, g_s = sub.logp_and_grad(st, add_priors=(rank == 0))
    lp_r, g_r = allreduce_logp_grad(torch.from_numpy(lp_s), torch.from_numpy(pack_grad(case, g_s)))
    err_obs = max(np.max(np.abs(lp_r.numpy() - lp_full) / np.abs(lp_full)),
                  np.max(np.abs(g_r.numpy() - g_full)) / np.max(np.abs(g_full)))

    # --- IWLS information of a location term: all-reduce of X'WX, then penalty / ridge /
    #     Cholesky on every rank (iwls_proposals.py:82-89) -----------------------------------
    x_s, _, _ = sub.loc_precision(st, 0)
    x_r = allreduce_sum(torch.from_numpy(x_s))
    P_r, L_r = finish_precision(x_r, torch.from_numpy(om.terms[0].K), torch.from_numpy(st.tau[:, 0]))
    _, P_full, L_full = om.loc_precision(st, 0)
    err_info = max(np.max(np.abs(P_r.numpy() - P_full)) / np.max(np.abs(P_full)),
                   np.max(np.abs(L_r.numpy() - L_full)) / np.max(np.abs(L_full)))

    # --- intercept pseudo-Gibbs statistics: three sums per chain, additive over shards -----
    stats = np.zeros((4, 5))
    for c in range(4):
        mu, ls = sub.predictors(st, c)
        w = np.exp(-ls)
        r = (sub.y - mu) * w
        stats[c] = [np.sum(w), np.sum(r), np.sum(r * r), np.sum(w * w), np.sum(r * w)]
    stats_r = allreduce_sum(torch.from_numpy(stats))
    params = torch.from_numpy(case.params)
    b0, g0 = intercepts_from_stats(stats_r, params, case.layout, om.n)
    ref = np.array([om.compute_intercepts(st, c) for c in range(4)])
    err_int = max(np.max(np.abs(b0.numpy() - ref[:, 0])), np.max(np.abs(g0.numpy() - ref[:, 1])))
    b0s, g0s = intercepts_from_stats(stats_r, params, case.layout, om.n, sequential=True)
    refs = np.array([om.compute_intercepts(st, c, sequential=True) for c in range(4)])
    err_int = max(err_int, np.max(np.abs(b0s.numpy() - refs[:, 0])), np.max(np.abs(g0s.numpy() - refs[:, 1])))
    err_obs = max(err_obs, err_info, err_int * 1e-1)  # intercepts to 1e-11 absolute

    # --- chain-axis sharding: no collective on the data path ---------------------------
    clo, chi = chain_shard(4, world, rank)
    sub_st = o.ChainState(beta=[b[clo:chi] for b in st.beta], tau=st.tau[clo:chi], delta_z=st.delta_z[clo:chi],
                          tau_delta=st.tau_delta[clo:chi], beta0=st.beta0[clo:chi], gamma0=st.gamma0[clo:chi])
    lp_c, _ = om.logp_and_grad(sub_st)
    gathered = [torch.zeros(chi - clo, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(lp_c))  # test-only gather to compare
    err_chain = float(np.max(np.abs(torch.cat(gathered).numpy() - lp_full)))
    q.put((rank, float(err_obs), err_chain))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_obs, err_chain in res:
        assert err_obs < 1e-12, (rank, err_obs)
        assert err_chain == 0.0, (rank, err_chain)
