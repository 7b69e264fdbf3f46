"""Non-centred terms (gam/var.py:176-197: value = basis . (scale * latent), latent ~ N(0,
inv(K)); IWLS factor scale * chol, iwls_proposals.py:82-89), the non-centred transformation
coefficient (var.py:100-107) and linear terms (gam/var.py:1025-1061) through the same kernels:
CUDA path through the C ABI against the oracle."""

from __future__ import annotations

import numpy as np
import pytest

import liesel_ptm_b200 as lp
from oracle import ptm_oracle as o
from tests.helpers import build_case, pack_grad, rel_err

pytestmark = pytest.mark.gpu


def _nc_case(dtype, trafo_nc, **shape):
    case = build_case(dtype=dtype, build=False, **shape)
    T = len(case.omodel.terms)
    for t in range(0, T, 2):  # every other term in non-centred form
        case.model.all_terms()[t][0].reparam_noncentered()
        case.omodel.terms[t].noncentered = True
    case.model.trafo_noncentered = trafo_nc
    case.omodel.trafo_noncentered = trafo_nc
    case.model.build()
    return case


@pytest.mark.parametrize("trafo_nc", [False, True])
def test_noncentered_logp_grad_information_f64(trafo_nc):
    import torch

    case = _nc_case(np.float64, trafo_nc, n=3000, n_loc=2, n_scale=2, nbases=12, C=7, amp=0.2)
    p = torch.from_numpy(case.params).cuda()
    lp_, g = case.model.log_prob_and_grad(p)
    lp_ref, gd = case.omodel.logp_and_grad(case.state)
    assert rel_err(lp_.cpu().numpy(), lp_ref) < 1e-12
    assert rel_err(case.model.log_prob(p).cpu().numpy(), lp_ref) < 1e-12
    g_ref = pack_grad(case, gd)
    for c in range(7):
        assert rel_err(g.cpu().numpy()[c], g_ref[c]) < 1e-12, c
    # shard views without priors keep the data part of the scale gradients
    a = case.model.shard_view(0, 1000, add_priors=True).log_prob_and_grad_block(p)
    b = case.model.shard_view(1000, 3000, add_priors=False).log_prob_and_grad_block(p)
    assert rel_err((a + b)[:, 1:].cpu().numpy(), g.cpu().numpy()) < 1e-11
    # IWLS information: the factor of a non-centred term is scale * chol
    for t in (0, 1):
        P, L = case.model.precision(t, p)
        _, P_ref, L_ref = case.omodel.loc_precision(case.state, t)
        assert rel_err(P.cpu().numpy(), P_ref) < 1e-11
        assert rel_err(L.cpu().numpy(), L_ref) < 1e-9
    # observed information of a non-centred block and of delta_z
    for blk in (0, 1, 2, "delta_z"):
        info, _ = case.model.hessian_block(blk, p, mode=0, with_chol=False)
        for c in (0, 6):
            assert rel_err(-info[c].cpu().numpy(), case.omodel.hessian_block(case.state, c, blk)) < 1e-11, blk


def test_noncentered_f32_tensor_core_route():
    import torch

    from tests.test_gpu_parity import _f64_oracle_of

    case = _nc_case(np.float32, True, n=5000, n_loc=2, n_scale=1, nbases=12, C=9, amp=0.15)
    p = torch.from_numpy(case.params).cuda()
    lp_, g = case.model.log_prob_and_grad(p)
    assert case.model.last_main_route() == "tcgen05"
    om, st = _f64_oracle_of(case)
    for t, term in enumerate(case.omodel.terms):
        om.terms[t].noncentered = term.noncentered
    om.trafo_noncentered = True
    lp_ref, gd = om.logp_and_grad(st)
    g_ref = pack_grad(case, gd)
    assert rel_err(lp_.cpu().numpy(), lp_ref) < 1e-5
    for c in range(9):
        assert rel_err(g.cpu().numpy()[c], g_ref[c]) < 1e-5, c


def test_linear_terms_vs_dense_linear_design():
    """``lin(x)`` and ``lin(x, add_intercept=True)``: the oracle uses the reference's dense design
    ``[x]`` / ``[1, x]`` (Basis.new_linear), the kernels the spline form with Greville
    coefficients."""
    import torch

    rng = np.random.default_rng(3)
    n, C = 2500, 5
    y, X = o.synth_data(n, 2, 1, seed=4)
    model = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=False, device="cuda:0")
    b0 = lp.lin(X[0], xname="x0")
    b1 = lp.ps(X[1], nbases=10, xname="x1")
    b2 = lp.lin(X[2], xname="x2", add_intercept=True)
    model.loc += lp.term.f_ig(b0, fname="lin")
    model.loc += lp.term.f_ig(b1, fname="f")
    model.scale += lp.term.f_ig(b2, fname="glin")
    designs = [X[0][:, None], None, np.stack([np.ones(n), X[2]], axis=1)]
    oterms = []
    for i, (b, pred) in enumerate(((b0, "loc"), (b1, "loc"), (b2, "scale"))):
        tm = o.term_from_reparam(X[i], b.knots, b.Z, b.penalty, int(np.linalg.matrix_rank(b.penalty)), pred)
        if designs[i] is not None:
            assert np.max(np.abs(tm.design - designs[i])) < 1e-14  # the spline form IS the linear design
            tm.design = designs[i]
        oterms.append(tm)
    om = o.PTMModel(y, oterms, nparam=20)
    om.Zd = lp.trafo_reparam(20)
    st = o.synth_state(oterms, 20, C, seed=5, amp=0.2)
    st.tau[:] = rng.uniform(0.6, 1.5, size=st.tau.shape)
    st.tau_delta[:] = rng.uniform(0.4, 0.9, size=C)
    params = model.pack(st.beta, st.delta_z, tau=st.tau, tau_delta=st.tau_delta, beta0=st.beta0, gamma0=st.gamma0)
    model.build()
    p = torch.from_numpy(params).cuda()
    lp_, g = model.log_prob_and_grad(p)
    lp_ref, gd = om.logp_and_grad(st)

    class _C:
        pass

    c_ = _C()
    c_.layout, c_.P = model.param_layout()
    g_ref = pack_grad(c_, gd)
    assert rel_err(lp_.cpu().numpy(), lp_ref) < 1e-12
    for c in range(C):
        assert rel_err(g.cpu().numpy()[c], g_ref[c]) < 1e-12
    P, L = model.precision(0, p)   # 1 x 1 information of the linear location term
    _, P_ref, L_ref = om.loc_precision(st, 0)
    assert rel_err(P.cpu().numpy(), P_ref) < 1e-11 and rel_err(L.cpu().numpy(), L_ref) < 1e-10
