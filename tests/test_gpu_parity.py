"""Parity of the CUDA path (through the C ABI) against the oracle.  f64: rel 1e-12
on logp / gradient (north_star); knot-interval indices bit-exact; basis values
bit-exact (same IEEE operation order, no FMA contraction)."""

import numpy as np
import pytest
import torch

from oracle import ptm_oracle as o
from tests.helpers import build_case, oracle_eval, rel_err

pytestmark = pytest.mark.gpu

TOL64 = 1e-12


def _eval(case):
    params = torch.from_numpy(case.params).to(case.model.device)
    logp, grad = case.model.log_prob_and_grad(params)
    logp_only = case.model.log_prob(params)
    torch.cuda.synchronize()
    return logp.cpu().numpy(), grad.cpu().numpy(), logp_only.cpu().numpy()


@pytest.mark.parametrize("n,C", [(1, 1), (7, 3), (100, 1), (500, 5), (4097, 3), (4097, 33)])
def test_logp_grad_f64(n, C):
    case = build_case(n=n, n_loc=2, n_scale=1, nbases=12, nparam=20, C=C)
    lp, g, lp_only = _eval(case)
    lp_ref, g_ref = oracle_eval(case)
    assert rel_err(lp, lp_ref) < TOL64
    assert rel_err(lp_only, lp_ref) < TOL64
    for c in range(C):
        assert rel_err(g[c], g_ref[c]) < TOL64, c


def test_basis_bit_exact():
    case = build_case(n=3000, n_loc=1, n_scale=1, nbases=20, C=1)
    for t, term in enumerate(case.omodel.terms):
        interval, values = case.model.basis(t)
        interval, values = interval.cpu().numpy(), values.cpu().numpy()
        D = o.basis_matrix(term.x, term.knots)
        assert np.array_equal(interval, np.searchsorted(term.knots, term.x, side="right") - 1)
        dense = np.zeros_like(D)
        for m in range(4):
            dense[np.arange(D.shape[0]), interval - 3 + m] = values[:, m]
        assert np.array_equal(dense, D)


def test_transform_matches_oracle():
    case = build_case(n=10, n_loc=1, n_scale=0, nbases=8, nparam=10, C=1)
    sp = case.omodel.spline
    rng = np.random.default_rng(1)
    delta = rng.normal(size=(3, 10))
    r = np.concatenate([np.linspace(-8, 8, 3001), sp.bspline.grid, [sp.min_eps, sp.max_eps]])
    z, dz = case.model.transform(torch.from_numpy(delta).cuda(), torch.from_numpy(r).cuda())
    z, dz = z.cpu().numpy(), dz.cpu().numpy()
    for c in range(3):
        zo, do = sp.dot_and_deriv_n_fullbatch(r, delta[c])
        assert np.max(np.abs(z[c] - zo)) < 1e-13 * 10
        assert np.max(np.abs(dz[c] - do)) < 1e-12
    cell = case.model.last_cell.cpu().numpy()
    core = sp.region_code(r) == 2
    ref_cell = np.searchsorted(sp.bspline.grid, r, side="right") - 1
    assert np.array_equal(cell[core], ref_cell[core])
    assert np.all(cell[~core] == -1)


def test_xtwx_and_precision():
    case = build_case(n=2000, n_loc=2, n_scale=2, nbases=12, C=4)
    params = torch.from_numpy(case.params).cuda()
    for t in range(2):
        xtwx = case.model.xtwx(t, params).cpu().numpy()
        P, L = case.model.precision(t, params)
        ref_x, ref_P, ref_L = case.omodel.loc_precision(case.state, t)
        assert rel_err(xtwx, ref_x) < 1e-11
        assert rel_err(P.cpu().numpy(), ref_P) < 1e-11
        assert rel_err(L.cpu().numpy(), ref_L) < 1e-9
    for t in (2, 3):
        P, _ = case.model.precision(t, params)
        _, ref_P = case.omodel.scale_precision(case.state, t)
        assert rel_err(P.cpu().numpy(), ref_P) < 1e-11
