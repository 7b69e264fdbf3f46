"""Observed information of a coefficient block:
``-jacfwd(grad(log_prob))`` as the reference's CholInfo family and kernel.IWLSKernel use it
(iwls_proposals.py:168-184, 227-245, 279-298, 333-364; iwls_fixed.py:118-143;
kernel.py:240-258).  CUDA path through the C ABI against (i) the fixture produced by the
reference's own code on nested forward-mode values and (ii) the oracle restatement."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import ptm_oracle as o
from tests import golden_io
from tests.helpers import build_case, rel_err

pytestmark = pytest.mark.gpu


def test_hessian_blocks_vs_reference_code_fixture():
    import torch

    from tests.test_reference_golden import _pack

    d = golden_io.load("ref_logp_n257.npz")
    n_h = int(d["hess_n"])
    model = golden_io.model_from_fixture({k: d[k] for k in d.files}, dtype=np.float64, n_rows=n_h).build()
    params = torch.from_numpy(_pack(model, d)).cuda()
    for blk, key in ((0, "hess_loc"), (1, "hess_scale"), ("delta_z", "hess_dz")):
        info, chol = model.hessian_block(blk, params[:1], mode=0)
        assert rel_err(-info[0].cpu().numpy(), d[key]) < 1e-11, key
        # the factor belongs to the symmetrised information (NaN where it is not positive
        # definite, like jnp.linalg.cholesky)
        assert rel_err(chol[0].cpu().numpy(), o.PTMModel.observed_info_chol(d[key], 0)[1]) < 1e-9, key


@pytest.mark.parametrize("shape", [dict(n=1500, n_loc=2, n_scale=2, nbases=12, C=5),
                                   dict(n=4097, n_loc=3, n_scale=1, nbases=20, C=33),
                                   dict(n=257, n_loc=1, n_scale=0, nbases=9, C=2)])
def test_hessian_blocks_vs_oracle_f64(shape):
    import torch

    # heavy tails so that transitions and linear tails carry weight in every block
    case = build_case(dtype=np.float64, amp=0.25, **shape)
    params = torch.from_numpy(case.params).cuda()
    T = len(case.omodel.terms)
    for blk in list(range(T)) + ["delta_z"]:
        info, chol = case.model.hessian_block(blk, params, mode=0)
        info = info.cpu().numpy()
        for c in range(0, shape["C"], max(1, shape["C"] // 4)):
            H = case.omodel.hessian_block(case.state, c, blk)
            assert rel_err(-info[c], H) < 1e-11, (blk, c)


def test_hessian_modes_match_reference_post_processing():
    import torch

    case = build_case(n=800, n_loc=2, n_scale=1, nbases=12, C=6, dtype=np.float64, amp=0.2)
    params = torch.from_numpy(case.params).cuda()
    for blk in (0, 2, "delta_z"):
        for mode in (0, 1, 2, 3):
            info, chol = case.model.hessian_block(blk, params, mode=mode)
            for c in (0, 5):
                H = case.omodel.hessian_block(case.state, c, blk)
                info_ref, L_ref = o.PTMModel.observed_info_chol(H, mode)
                if mode == 3:  # kernel.IWLSKernel: the factor is of info + I, the matrix stays raw
                    info_ref = -H
                assert rel_err(info[c].cpu().numpy(), info_ref) < 1e-11, (blk, mode)
                # factor: the last pivots of an ill-conditioned block amplify rounding, so
                # the factor itself to 1e-7 and its product to 1e-11
                Lc = chol[c].cpu().numpy()
                assert rel_err(Lc, L_ref) < 1e-7, (blk, mode)
                assert rel_err(Lc @ Lc.T, L_ref @ L_ref.T) < 1e-11, (blk, mode)


def test_hessian_make_pd_shifts_an_indefinite_block():
    """A location block whose information is rank-deficient (no penalty: tau = inf is not
    representable, so use a shard view without priors is refused; instead n < p observations
    and a huge tau) needs the eigen-shift of make_pd (iwls_proposals.py:236-241)."""
    import torch

    case = build_case(n=6, n_loc=1, n_scale=0, nbases=14, C=3, dtype=np.float64, tau=1e8)
    params = torch.from_numpy(case.params).cuda()
    raw, L0 = case.model.hessian_block(0, params, mode=0)
    pd, L1 = case.model.hessian_block(0, params, mode=1)
    for c in range(3):
        H = case.omodel.hessian_block(case.state, c, 0)
        info_ref, L_ref = o.PTMModel.observed_info_chol(H, 1)
        w = np.linalg.eigvalsh(0.5 * (-H - H.T))
        assert w.min() < 1e-6  # the case is what it says
        assert rel_err(pd[c].cpu().numpy(), info_ref) < 1e-9
        assert np.linalg.eigvalsh(pd[c].cpu().numpy()).min() > 0.5e-6
        assert np.isfinite(L1[c].cpu().numpy()).all()
    # mode 4: identity when the factor of info + 1e-6 I has NaNs, else that factor
    _, L4 = case.model.hessian_block(0, params, mode=4)
    L4 = L4.cpu().numpy()
    for c in range(3):
        H = case.omodel.hessian_block(case.state, c, 0)
        A = -H + 1e-6 * np.eye(H.shape[0])
        try:
            ref = np.linalg.cholesky(0.5 * (A + A.T))
            if np.isnan(ref).any():
                raise np.linalg.LinAlgError
            # near-singular: compare the product, not the factor
            assert rel_err(L4[c] @ L4[c].T, 0.5 * (A + A.T)) < 1e-8
        except np.linalg.LinAlgError:
            assert np.array_equal(L4[c], np.eye(H.shape[0]))


def test_hessian_nan_response_and_shards():
    import torch

    case = build_case(n=900, n_loc=1, n_scale=1, nbases=10, C=4, dtype=np.float64)
    params = torch.from_numpy(case.params).cuda()
    full, _ = case.model.hessian_block(0, params, mode=0, with_chol=False)
    a = case.model.shard_view(0, 400, add_priors=True).hessian_block(0, params, mode=0, with_chol=False)[0]
    b = case.model.shard_view(400, 900, add_priors=False).hessian_block(0, params, mode=0, with_chol=False)[0]
    assert rel_err((a + b).cpu().numpy(), full.cpu().numpy()) < 1e-12
    with pytest.raises(Exception):
        case.model.shard_view(0, 400).hessian_block(0, params, mode=1)
    bad = build_case(n=300, n_loc=1, n_scale=1, nbases=10, C=2, dtype=np.float64,
                     y_override=lambda y: np.where(np.arange(y.shape[0]) == 17, np.nan, y))
    for blk in (0, 1, "delta_z"):
        info, _ = bad.model.hessian_block(blk, torch.from_numpy(bad.params).cuda())
        assert np.isnan(info.cpu().numpy()).any(), blk


def test_hessian_f32_vs_f64_oracle():
    import torch

    case = build_case(n=2000, n_loc=2, n_scale=1, nbases=12, C=4, dtype=np.float32, amp=0.15)
    params = torch.from_numpy(case.params).cuda()
    # float64 oracle on the float32-rounded inputs
    case64 = build_case(n=2000, n_loc=2, n_scale=1, nbases=12, C=4, dtype=np.float32, amp=0.15, build=False)
    om = case64.omodel
    for blk in (0, 2):
        info, _ = case.model.hessian_block(blk, params, with_chol=False)
        for c in (0, 3):
            H = om.hessian_block(case64.state, c, blk)
            # float32 oracle arithmetic vs float kernel with double accumulation: 1e-4 of the norm
            assert rel_err(-info[c].cpu().numpy(), H) < 2e-4, blk


def test_observed_chol_info_providers():
    import torch

    import liesel_ptm_b200.model as M

    case = build_case(n=600, n_loc=1, n_scale=1, nbases=10, C=3, dtype=np.float64)
    params = torch.from_numpy(case.params).cuda()
    ci = M.CholInfo(case.model, 0, params)
    H = case.omodel.hessian_block(case.state, 1, 0)
    info_ref, L_ref = o.PTMModel.observed_info_chol(H, 2)
    assert rel_err(ci.fisher_info[1].cpu().numpy(), info_ref) < 1e-11
    assert rel_err(ci.chol_info()[1].cpu().numpy(), L_ref) < 1e-7
    raw_fails = any(np.isnan(o.PTMModel.observed_info_chol(case.omodel.hessian_block(case.state, c, 0), 0)[1]).any()
                    for c in range(3))
    assert ci.nan_in_cholesky_of_unprocessed_finfo == raw_fails
    pf = M.PTMCholInfoFixed.from_coef(case.model, params)
    Hd = case.omodel.hessian_block(case.state, 2, "delta_z")
    assert rel_err(pf.chol_info()[2].cpu().numpy(), o.PTMModel.observed_info_chol(Hd, 1)[1]) < 1e-7
    pc = M.PTMCholInfo.from_coef(case.model, params)
    td = case.state.tau_delta[2]
    P_ref = o.PTMModel.observed_info_chol(Hd, 1)[0] / td**2
    assert rel_err(pc.precision(params)[2].cpu().numpy(), P_ref) < 1e-11
    L3 = M.iwls_kernel_chol_info(case.model, 0, params)
    assert rel_err(L3[0].cpu().numpy(),
                   o.PTMModel.observed_info_chol(case.omodel.hessian_block(case.state, 0, 0), 3)[1]) < 1e-7
