"""Tensor-core route of the float32 logp(+gradient) sweep (ptm_main_tc.cuh: eta = Gamma B' and
the coefficient gradients G' B on tcgen05 with the 3xTF32 split) against the float64 oracle on
the float32-rounded inputs -- north star: rel 1e-5 -- and against the CUDA-core route.
Reference: dist.py:185-188, 339-395, 718-740; gam/var.py:157-162; logprob.py:104-113."""

from __future__ import annotations

import numpy as np
import pytest

from tests import golden_io
from tests.helpers import build_case, pack_grad_layout, rel_err

pytestmark = pytest.mark.gpu


def _oracle64(case):
    """float64 oracle evaluated on the float32-rounded numbers the kernel receives."""
    from tests.test_gpu_parity import _f64_oracle_of

    return _f64_oracle_of(case)


SHAPES = [
    dict(n=4097, n_loc=2, n_scale=2, nbases=12, C=70),      # K = 48: 64-observation tiles
    dict(n=20011, n_loc=5, n_scale=5, nbases=20, C=130),    # the c3 term structure: K = 208
    dict(n=333, n_loc=1, n_scale=0, nbases=9, C=3),         # location terms only
    dict(n=50, n_loc=0, n_scale=1, nbases=20, C=257),       # scale terms only, three chain blocks
    dict(n=1, n_loc=1, n_scale=1, nbases=8, C=1),
    dict(n=6000, n_loc=4, n_scale=3, nbases=24, C=33),      # K = 168: 48-observation tiles
    dict(n=7000, n_loc=10, n_scale=10, nbases=20, C=40),    # the c5 term structure: K = 200 + 200, two-pass form
    dict(n=3000, n_loc=7, n_scale=2, nbases=30, C=129),     # K = 224 + 64: two-pass, unequal blocks
    dict(n=2500, n_loc=10, n_scale=1, nbases=30, C=20),     # the c4 term structure: 320 + 32, location split
    dict(n=2000, n_loc=16, n_scale=6, nbases=30, C=5),      # 512 + 192: three eta-only launches, accumulation
]


@pytest.mark.parametrize("shape", SHAPES)
def test_tc_route_logp_grad_vs_f64_oracle(shape):
    import torch

    case = build_case(dtype=np.float32, amp=0.15, **shape)
    p = torch.from_numpy(case.params).cuda()
    lp, g = case.model.log_prob_and_grad(p)
    assert case.model.last_main_route() == "tcgen05"
    lp_only = case.model.log_prob(p)
    assert case.model.last_main_route() == "tcgen05"
    block = case.model.log_prob_and_grad_block(p)
    om, st = _oracle64(case)
    lp_ref, gd = om.logp_and_grad(st)
    g_ref = pack_grad_layout(case.layout, case.P, gd)
    assert rel_err(lp.cpu().numpy(), lp_ref) < 1e-5
    assert rel_err(lp_only.cpu().numpy(), lp_ref) < 1e-5
    gn = g.cpu().numpy()
    for c in range(gn.shape[0]):
        assert rel_err(gn[c], g_ref[c]) < 1e-5, c
    assert np.array_equal(block[:, 0].cpu().numpy(), lp.cpu().numpy())
    assert np.array_equal(block[:, 1:].cpu().numpy(), gn)
    # deterministic: fixed reduction order, single-writer promotions
    lp2, g2 = case.model.log_prob_and_grad(p)
    assert torch.equal(lp2, lp) and torch.equal(g2, g)


def test_tc_route_selection(monkeypatch):
    import torch

    small = build_case(n=700, n_loc=2, n_scale=1, nbases=12, C=4, dtype=np.float32)
    p = torch.from_numpy(small.params).cuda()
    small.model.log_prob(p)
    assert small.model.last_main_route() == "tcgen05"
    # float64 problems and models whose coefficients do not fit tensor memory: CUDA-core sweep
    f64 = build_case(n=700, n_loc=2, n_scale=1, nbases=12, C=4, dtype=np.float64)
    f64.model.log_prob(torch.from_numpy(f64.params).cuda())
    assert f64.model.last_main_route() == "cuda-core"
    f64.model.log_prob_and_grad(torch.from_numpy(f64.params).cuda())
    assert f64.model.last_main_route() == "cuda-core"
    # the knob is read when the problem is created
    monkeypatch.setenv("PTM_MAIN_TC", "0")
    off = build_case(n=700, n_loc=2, n_scale=1, nbases=12, C=4, dtype=np.float32)
    lp0, g0 = off.model.log_prob_and_grad(p)
    assert off.model.last_main_route() == "cuda-core"
    lp1, g1 = small.model.log_prob_and_grad(p)
    assert small.model.last_main_route() == "tcgen05"
    assert rel_err(lp1.cpu().numpy(), lp0.cpu().numpy()) < 2e-6
    for c in range(4):
        assert rel_err(g1[c].cpu().numpy(), g0[c].cpu().numpy()) < 1e-5


def test_tc_route_nan_shards_and_intercept_stats():
    import torch

    case = build_case(n=5000, n_loc=2, n_scale=2, nbases=12, C=40, dtype=np.float32)
    p = torch.from_numpy(case.params).cuda()
    lp, g = case.model.log_prob_and_grad(p)
    # observation shards add up (priors once): the packed block of the observation-axis all-reduce
    acc = None
    for lo, hi, pri in ((0, 1, True), (1, 1777, False), (1777, 5000, False)):
        b = case.model.shard_view(lo, hi, add_priors=pri).log_prob_and_grad_block(p)
        acc = b.double() if acc is None else acc + b.double()
    assert rel_err(acc[:, 0].cpu().numpy(), lp.cpu().numpy()) < 2e-6
    for c in range(0, 40, 7):
        assert rel_err(acc[c, 1:].cpu().numpy(), g[c].cpu().numpy()) < 1e-5
    # intercept statistics ride along with the logp-only sweep (predictor.py:80-100, 126-130)
    om, st = _oracle64(case)
    b0, g0 = case.model.compute_intercepts(p)
    for c in (0, 13, 39):
        rb, rg = om.compute_intercepts(st, c)
        assert abs(float(b0[c]) - rb) < 1e-5 and abs(float(g0[c]) - rg) < 1e-5
    # NaN response -> NaN logp (dist.py:340-343), every chain
    bad = build_case(n=900, n_loc=1, n_scale=1, nbases=10, C=5, dtype=np.float32,
                     y_override=lambda y: np.where(np.arange(y.shape[0]) == 123, np.nan, y))
    lpn = bad.model.log_prob(torch.from_numpy(bad.params).cuda())
    assert bad.model.last_main_route() == "tcgen05"
    assert np.isnan(lpn.cpu().numpy()).all()


def test_tc_route_matches_cuda_core_route_at_c3_size(monkeypatch):
    """n = 1M, the c3 term structure, 128 chains: both float routes against each other (logp
    2e-6, gradient 2e-5 of its inf-norm: each route is within ~7e-6 / 2e-6 of the float64 kernel,
    tests/test_gpu_parity.py::test_full_size_f32_vs_f64 measures that bound)."""
    import torch

    import bench_workload as bw

    wl = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=128, dtype=np.float32)
    m = wl.model.build()
    p = torch.from_numpy(wl.params).cuda()
    lp, g = m.log_prob_and_grad(p)
    assert m.last_main_route() == "tcgen05"
    monkeypatch.setenv("PTM_MAIN_TC", "0")
    wl0 = bw.build("c3_n1M_5loc_5scale_nb20_1024chains", chains=128, dtype=np.float32)
    m0 = wl0.model.build()
    lp0, g0 = m0.log_prob_and_grad(p)
    assert m0.last_main_route() == "cuda-core"
    assert rel_err(lp.cpu().numpy(), lp0.cpu().numpy()) < 2e-6
    e = ((g.double() - g0.double()).abs().amax(dim=1) / g0.double().abs().amax(dim=1)).max().item()
    assert e < 2e-5, e


def test_tc_route_random_shapes_fuzz():
    """tools/fuzz_tc.py: 40 seeded random (n, chains, terms, nbases, transformation variant) models,
    several-pass forms included -- the tensor-core route against the CUDA-core route (logp 5e-6,
    gradient 3e-5 of its inf-norm) and, for the small ones, the float64 oracle at 1e-5."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_tc.py"), "40", "11"], cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "failures: 0" in r.stdout, r.stdout[-3000:]
