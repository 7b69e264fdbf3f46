"""The other two transformation variants PTMDist switches (model/model.py:98-131): the onion
spline (bspline/onion.py:13-138) and the Gaussian model (GaussianPseudoTransformationDist,
dist.py:766-850; LocScalePTM.new_gaussian model/model.py:386-417).

tests/golden/ref_variants_n257.npz holds what the reference's own source files produce for them
on the data and chain states of ref_logp_n257.npz (tests/golden/make_reference_golden.py,
``ref_variants``): knots, theta(delta), h / h' inside and outside the core, per-observation
log-densities, the forward-mode gradient with the declared custom_jvp rules, and the
jacfwd(grad) blocks.  CPU part: the oracle against them.  GPU part: the CUDA path, through the
C ABI, against them and against the oracle."""

from __future__ import annotations

import numpy as np
import pytest

import liesel_ptm_b200 as lp
from oracle import ptm_oracle as o
from tests import golden_io
from tests.helpers import pack_grad_layout, rel_err


def _fixtures():
    return golden_io.load("ref_logp_n257.npz"), golden_io.load("ref_variants_n257.npz")


def _state(d, C=2):
    p = d["Z0"].shape[1]
    P = d["params"]
    return o.ChainState(beta=[P[:, :p], P[:, p:2 * p]], tau=d["tau"], delta_z=P[:, 2 * p:],
                        tau_delta=d["tau_delta"], beta0=np.zeros(C), gamma0=np.zeros(C))


def _flat_grad(d, gd):
    return np.concatenate([gd["beta"][0], gd["beta"][1], gd["delta_z"]], axis=1)


# ---------------------------------------------------------------------------------------
# oracle vs the reference's code (CPU)
# ---------------------------------------------------------------------------------------
def test_oracle_onion_knots_coef_and_spline_match_reference_code():
    _, v = _fixtures()
    kn = o.OnionKnots(-4.0, 4.0, 20)
    assert np.array_equal(kn.knots, v["onion_knots"]) and kn.step == v["onion_step"]
    assert np.array_equal(lp.OnionKnots(-4.0, 4.0, 20).knots, v["onion_knots"])
    sp = o.OnionSpline(kn.knots)
    assert sp.min_knot == v["onion_min_knot"] and sp.max_knot == v["onion_max_knot"]
    assert np.array_equal(sp.bspline.grid, v["onion_grid"])
    for c in range(3):
        theta = sp.compute_coef(v["delta"][c])
        assert theta.shape == (27,)
        assert rel_err(theta, v["onion_theta"][c]) < 1e-15
        z, d = sp.dot_and_deriv_n_fullbatch(v["r"], v["delta"][c])
        assert np.max(np.abs(z - v["onion_z"][c])) < 1e-14
        assert np.max(np.abs(d - v["onion_d"][c])) < 1e-13
    # identity off the core, whatever the coefficients (onion.py:133-138)
    out = (v["r"] < sp.min_knot) | (v["r"] > sp.max_knot)
    assert out.any()
    assert np.array_equal(v["onion_z"][1][out], v["r"][out]) and np.all(v["onion_d"][1][out] == 1.0)
    # delta = 0: all increments equal the knot step => the identity on the core too, up to the
    # lerp weight of the reference's table (grid spacing range/999, weight step range/1000,
    # approx.py:375-380, 425-428: 0.1 % of a cell)
    z0, d0 = sp.dot_and_deriv_n_fullbatch(v["r"], v["delta"][0])
    assert np.max(np.abs(z0 - v["r"])) < 1e-5 and np.max(np.abs(d0 - 1.0)) < 1e-12


def test_oracle_onion_properties_of_the_reference_tests():
    """tests/test_bspline/test_onion.py of the reference: OnionKnots(-4, 4, nparam=10), normal
    coefficients, x on linspace(-8, 8): no NaN, h' > 0, h increasing."""
    kn = o.OnionKnots(-4.0, 4.0, 10)
    sp = o.OnionSpline(kn.knots)
    rng = np.random.default_rng(1)
    for x in (np.linspace(-8.0, 8.0, 300), np.linspace(-8.0, 8.0, 10_000)):
        for coef in rng.normal(size=(3, 10)):
            fx, fxd = sp.dot_and_deriv_n_fullbatch(x, coef)
            assert fx.shape == x.shape and fxd.shape == x.shape
            assert not np.isnan(fx).any() and not np.isnan(fxd).any()
            assert np.all(fxd > 0.0) and np.all(np.diff(fx) > 0.0)


@pytest.mark.parametrize("kind", ["onion", "identity"])
def test_oracle_variant_logp_grad_match_reference_code(kind):
    d, v = _fixtures()
    om = golden_io.oracle_from_fixture(d, bspline=kind)
    st = _state(d)
    for c in range(2):
        assert rel_err(om.loglik_terms(st, c)["ll"], v[f"{kind}_loglik_i"][c]) < 1e-13
    lp_, gd = om.logp_and_grad(st)
    assert rel_err(lp_, v[f"{kind}_logp"]) < 1e-13
    g_ref = v[f"{kind}_grad"]
    assert rel_err(_flat_grad(d, gd), g_ref) < 1e-12
    if kind == "identity":  # no shape parameters in the Gaussian model
        p = d["Z0"].shape[1]
        assert np.all(g_ref[:, 2 * p:] == 0.0) and np.all(gd["delta_z"] == 0.0) and np.all(gd["tau_delta"] == 0.0)


def test_oracle_variant_hessians_match_reference_code():
    d, v = _fixtures()
    rows = slice(0, int(v["hess_n"]))
    st = _state(d)
    om = golden_io.oracle_from_fixture(d, bspline="onion")
    for block, key in ((0, "loc"), (1, "scale"), ("delta_z", "dz")):
        H = om.hessian_block(st, 0, block, rows=rows)
        ref = v[f"hess_onion_{key}"]
        assert np.max(np.abs(H - ref)) < 1e-11 * np.max(np.abs(ref)), key
    og = golden_io.oracle_from_fixture(d, bspline="identity")
    for block, key in ((0, "loc"), (1, "scale")):
        H = og.hessian_block(st, 0, block, rows=rows)
        ref = v[f"hess_identity_{key}"]
        assert np.max(np.abs(H - ref)) < 1e-11 * np.max(np.abs(ref)), key
    with pytest.raises(ValueError):
        og.hessian_block(st, 0, "delta_z", rows=rows)


# ---------------------------------------------------------------------------------------
# the CUDA path (GPU, through the C ABI)
# ---------------------------------------------------------------------------------------
def _gpu_model(d, kind, dtype=np.float64, n_rows=None):
    dd = {k: d[k] for k in d.files}
    return golden_io.model_from_fixture(dd, dtype=dtype, bspline=kind, n_rows=n_rows).build()


def _pack(model, d, dtype=np.float64):
    p = d["Z0"].shape[1]
    P = d["params"]
    return model.pack([P[:, :p], P[:, p:2 * p]], P[:, 2 * p:], tau=d["tau"], tau_delta=d["tau_delta"]).astype(dtype)


def _cols(model, d):
    L, _ = model.param_layout()
    p = d["Z0"].shape[1]
    return np.r_[L["beta"][0]:L["beta"][0] + p, L["beta"][1]:L["beta"][1] + p,
                 L["delta_z"]:L["delta_z"] + d["params"].shape[1] - 2 * p]


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["onion", "identity"])
def test_cuda_variant_logp_grad_vs_reference_code_f64(kind):
    import torch

    d, v = _fixtures()
    model = _gpu_model(d, kind)
    params = torch.from_numpy(_pack(model, d)).cuda()
    logp, grad = model.log_prob_and_grad(params)
    assert rel_err(logp.cpu().numpy(), v[f"{kind}_logp"]) < 1e-12
    assert rel_err(model.log_prob(params).cpu().numpy(), v[f"{kind}_logp"]) < 1e-12
    g = grad.cpu().numpy()
    assert rel_err(g[:, _cols(model, d)], v[f"{kind}_grad"]) < 1e-12
    # every parameter (intercepts, smoothing scales too) against the oracle
    om = golden_io.oracle_from_fixture(d, bspline=kind)
    L, P = model.param_layout()
    st = golden_io.state_from_params(om, params.cpu().numpy(), L)
    lp_ref, gd = om.logp_and_grad(st)
    g_ref = pack_grad_layout(L, P, gd)
    for c in range(2):
        assert rel_err(g[c], g_ref[c]) < 1e-12
    # per-observation log-densities through the pointwise sweep
    ll = v[f"{kind}_loglik_i"]
    mx = ll.max(axis=0)
    lppd_ref = mx + np.log(np.exp(ll - mx).sum(axis=0)) - np.log(ll.shape[0])
    lppd, _ = model.pointwise(params)
    assert rel_err(lppd.cpu().numpy(), lppd_ref) < 1e-12
    # observation shards add up
    acc = None
    for lo, hi, pri in ((0, 100, True), (100, 257, False)):
        b = model.shard_view(lo, hi, add_priors=pri).log_prob_and_grad_block(params)
        acc = b if acc is None else acc + b
    assert rel_err(acc[:, 0].cpu().numpy(), v[f"{kind}_logp"]) < 1e-12
    if kind == "identity":
        assert np.all(g[:, L["delta_z"]:L["delta_z"] + 19] == 0.0) and np.all(g[:, L["tau_delta"]] == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["onion", "identity"])
def test_cuda_variant_logp_grad_f32_both_routes(kind, monkeypatch):
    """float32 problems: rel 1e-5 (north star) against the float64 oracle on the float32-rounded
    inputs, on the tensor-core route and on the CUDA-core route."""
    import torch

    d, _ = _fixtures()
    d32 = {k: (d[k].astype(np.float32).astype(np.float64) if d[k].dtype == np.float64 and d[k].ndim > 0
               else d[k]) for k in d.files}
    for tc in ("1", "0"):
        monkeypatch.setenv("PTM_MAIN_TC", tc)
        model = _gpu_model(d, kind, dtype=np.float32)
        p32 = _pack(model, d, np.float32)
        params = torch.from_numpy(p32).cuda()
        logp, grad = model.log_prob_and_grad(params)
        assert model.last_main_route() == ("tcgen05" if tc == "1" else "cuda-core")
        om = golden_io.oracle_from_fixture(d32, bspline=kind)
        if kind == "onion":
            om.knots = o.OnionKnots(-4.0, 4.0, 20, dtype=np.float32)
            om.spline = o.OnionSpline(om.knots.knots.astype(np.float64))
        L, P = model.param_layout()
        st = golden_io.state_from_params(om, p32.astype(np.float64), L)
        lp_ref, gd = om.logp_and_grad(st)
        g_ref = pack_grad_layout(L, P, gd)
        assert rel_err(logp.cpu().numpy(), lp_ref) < 1e-5
        g = grad.cpu().numpy()
        for c in range(2):
            assert rel_err(g[c], g_ref[c]) < 1e-5, (tc, c)


@pytest.mark.gpu
def test_cuda_onion_transform_vs_reference_code_and_properties():
    """ptm_transform_* / ptm_inverse_* on an onion problem: h, h' against the reference-run fixture
    inside and outside the core, the properties the reference's test_onion.py asserts, and the
    round trip through the inverse."""
    import torch

    _, v = _fixtures()
    y = np.zeros(8)
    model = lp.LocScalePTM(y, lp.OnionKnots(-4.0, 4.0, 20).knots, to_float32=False, device="cuda:0",
                           bspline="onion").build()
    z, dz = model.transform(torch.from_numpy(v["delta"]).cuda(), torch.from_numpy(v["r"]).cuda())
    assert np.max(np.abs(z.cpu().numpy() - v["onion_z"])) < 2e-13
    assert np.max(np.abs(dz.cpu().numpy() - v["onion_d"])) < 2e-12
    # nparam = 10 like the reference's tests
    m10 = lp.LocScalePTM(y, lp.OnionKnots(-4.0, 4.0, 10).knots, to_float32=False, device="cuda:0",
                         bspline="onion").build()
    rng = np.random.default_rng(1)
    coef = torch.from_numpy(rng.normal(size=(3, 10))).cuda()
    x = torch.linspace(-8.0, 8.0, 10_000, dtype=torch.float64, device="cuda")
    fx, fxd = m10.transform(coef, x)
    assert fx.shape == (3, 10_000) and fxd.shape == (3, 10_000)
    assert not torch.isnan(fx).any() and not torch.isnan(fxd).any()
    assert bool((fxd > 0.0).all()) and bool((fx[:, 1:] - fx[:, :-1] > 0.0).all())
    sp = o.OnionSpline(o.OnionKnots(-4.0, 4.0, 10).knots)
    for c in range(3):
        rz, rd = sp.dot_and_deriv_n_fullbatch(x.cpu().numpy(), coef[c].cpu().numpy())
        assert np.max(np.abs(fx[c].cpu().numpy() - rz)) < 1e-12 and np.max(np.abs(fxd[c].cpu().numpy() - rd)) < 1e-11
    # inverse: h(h^-1(z)) = z (the reference's round-trip tests accept 1e-5 .. 1e-4,
    # tests/test_bspline/test_ptm.py:167-301)
    zq = torch.linspace(-7.0, 7.0, 4001, dtype=torch.float64, device="cuda")
    for c in range(3):
        r_inv = m10.inverse(coef[c], zq)
        z_back, _ = m10.transform(coef[c], r_inv)
        assert float((z_back - zq).abs().max()) < 1e-9


@pytest.mark.gpu
def test_cuda_variant_hessians_vs_reference_code():
    import torch

    d, v = _fixtures()
    n_h = int(v["hess_n"])
    for kind, blocks in (("onion", ((0, "loc"), (1, "scale"), ("delta_z", "dz"))),
                         ("identity", ((0, "loc"), (1, "scale")))):
        model = _gpu_model(d, kind, n_rows=n_h)
        params = torch.from_numpy(_pack(model, d)).cuda()
        for block, key in blocks:
            info, _ = model.hessian_block(block, params[:1], mode=0)
            ref = -v[f"hess_{kind}_{key}"]
            assert np.max(np.abs(info[0].cpu().numpy() - ref)) < 1e-11 * np.max(np.abs(ref)), (kind, key)
        if kind == "identity":
            with pytest.raises(RuntimeError):
                model.hessian_block("delta_z", params[:1], mode=0)


@pytest.mark.gpu
def test_cuda_new_gaussian_and_onion_at_size():
    """new_gaussian (model/model.py:386-417) against the closed form of the Gaussian
    location-scale log-likelihood, and an onion model at n = 50k against the oracle."""
    import torch

    from tests.helpers import build_case

    rng = np.random.default_rng(3)
    n = 4000
    x = rng.uniform(-2, 2, size=n)
    y = np.sin(x) + 0.5 * rng.normal(size=n)
    model = lp.LocScalePTM.new_gaussian(y, to_float32=False, device="cuda:0")
    model.loc += lp.term.f_ig(lp.ps(x, nbases=10, xname="x"), fname="f")
    model.scale += lp.term.f_ig(lp.ps(x, nbases=8, xname="x"), fname="g")
    model.build()
    L, P = model.param_layout()
    params = np.zeros((3, P))
    params[:, L["beta0"]] = [0.0, 0.1, -0.2]
    params[:, L["gamma0"]] = [-0.7, -0.5, 0.2]
    for t in range(2):
        pt = model.terms[t][0].basis.ncoef
        params[:, L["beta"][t]:L["beta"][t] + pt] = 0.3 * rng.normal(size=(3, pt))
        params[:, L["tau"][t]] = [0.7, 1.0, 1.4]
    params[:, L["tau_delta"]] = 1.0
    params[:, L["delta_z"]:L["delta_z"] + model.nparam - 1] = rng.normal(size=(3, model.nparam - 1))  # ignored
    pt_ = torch.from_numpy(params).cuda()
    logp, grad = model.log_prob_and_grad(pt_)
    ref = np.zeros(3)
    for c in range(3):
        mu = np.full(n, params[c, L["beta0"]])
        ls = np.full(n, params[c, L["gamma0"]])
        for t, (tm, kind) in enumerate(model.terms):
            j, vals = lp.banded_basis(x, tm.basis.knots)
            dense = np.zeros((n, tm.basis.nbases + 1))
            for m in range(4):
                dense[np.arange(n), j - 3 + m] = vals[:, m]
            f = dense[:, :tm.basis.nbases] @ (tm.basis.Z @ params[c, L["beta"][t]:L["beta"][t] + tm.basis.ncoef])
            if kind == 0:
                mu += f
            else:
                ls += f
            b = params[c, L["beta"][t]:L["beta"][t] + tm.basis.ncoef]
            tau = params[c, L["tau"][t]]
            ref[c] += -(tm.penalty_rank * np.log(tau) + 0.5 * b @ tm.penalty @ b / tau**2)
        r = (y - mu) * np.exp(-ls)
        ref[c] += np.sum(-0.5 * r * r - 0.5 * np.log(2 * np.pi) - ls)
    assert rel_err(logp.cpu().numpy(), ref) < 1e-12
    g = grad.cpu().numpy()
    assert np.all(g[:, L["delta_z"]:L["delta_z"] + model.nparam - 1] == 0.0)

    case = build_case(n=50_000, n_loc=2, n_scale=2, nbases=12, C=5, dtype=np.float64, bspline="onion")
    p = torch.from_numpy(case.params).cuda()
    lp_, g_ = case.model.log_prob_and_grad(p)
    lp_ref, gd = case.omodel.logp_and_grad(case.state)
    g_ref = pack_grad_layout(case.layout, case.P, gd)
    assert rel_err(lp_.cpu().numpy(), lp_ref) < 1e-12
    for c in range(5):
        assert rel_err(g_[c].cpu().numpy(), g_ref[c]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["onion", "identity"])
def test_cuda_variant_predictive_records_vs_oracle(kind):
    """init_dist(samples, newdata=...) -> transformation / log_prob / cdf (model/model.py:698-790) and
    the pointwise WAIC records of the training response for the onion and the Gaussian model."""
    import torch

    from tests.helpers import build_case

    case = build_case(n=400, n_loc=2, n_scale=2, nbases=10, nparam=20, C=7, amp=0.3, bspline=kind)
    rng = np.random.default_rng(31)
    m = 333
    newdata = {f"x{t}": rng.uniform(-1.99, 1.99, size=m) for t in range(4)}
    grid = rng.normal(scale=2.5, size=m)
    params = torch.from_numpy(case.params).cuda()
    out = case.model.summarise_dist(params, grid=grid, newdata=newdata)
    X = [newdata[f"x{t}"] for t in range(4)]
    for c in range(7):
        z, lp_, cdf, _, _ = case.omodel.predict(case.state, c, grid, X)
        assert rel_err(out["z"][c].cpu().numpy(), z) < 1e-12
        assert rel_err(out["log_prob"][c].cpu().numpy(), lp_) < 1e-12
        assert rel_err(out["cdf"][c].cpu().numpy(), cdf) < 1e-12
    ll = np.stack([case.omodel.loglik_terms(case.state, c)["ll"] for c in range(7)])
    mx = ll.max(axis=0)
    lppd_ref = mx + np.log(np.exp(ll - mx).sum(axis=0)) - np.log(7)
    lppd, pwaic = case.model.pointwise(params)
    assert rel_err(lppd.cpu().numpy(), lppd_ref) < 1e-12
    assert rel_err(pwaic.cpu().numpy(), ll.var(axis=0)) < 1e-9
    # quantile -> cdf round trip through the CUDA path
    u = rng.uniform(0.01, 0.99, size=m)
    yq = case.model.predict_quantile(params, u, newdata=newdata)
    for c in range(2):
        cdf = case.model.summarise_dist(params[c:c + 1], grid=yq[c].cpu().numpy(), newdata=newdata, which=("cdf",))["cdf"]
        assert np.max(np.abs(cdf.cpu().numpy()[0] - u)) < 1e-10


@pytest.mark.gpu
def test_cuda_gaussian_model_sampler_recovers_the_simulated_truth():
    """LocScalePTM.new_gaussian (model/model.py:386-417) through the sampler composition of
    liesel_ptm_b200/mcmc.py (IWLS per smooth term, tau^2 Gibbs, pseudo-Gibbs intercepts; no shape
    block): the posterior mean of f(x) recovers the simulated function and the residual scale."""
    import torch

    from liesel_ptm_b200 import mcmc

    rng = np.random.default_rng(5)
    n, C = 1500, 16
    x = rng.uniform(-2, 2, n)
    f_true = 0.8 * np.sin(1.5 * x)
    y = 0.3 + f_true + 0.5 * rng.normal(size=n)
    model = lp.LocScalePTM.new_gaussian(y, to_float32=False, device="cuda:0")
    basis = lp.ps(x, nbases=12, xname="x")
    model.loc += lp.term.f_ig(basis, fname="f")
    model.build()
    L = model.layout
    p = basis.ncoef
    start = model.pack([rng.normal(0, 0.01, (C, p))], np.zeros((C, model.nparam - 1)), tau=1.0, tau_delta=1.0,
                       beta0=np.full(C, y.mean()), gamma0=np.full(C, np.log(y.std())))
    dr, info = mcmc.run_chains(model, torch.from_numpy(start).cuda(), warmup=150, draws=150, seed=2,
                               sample_tau=True, pgibbs_intercepts=True, use_graph=False)
    assert np.isfinite(dr).all()
    # the ignored shape entries stay where they were put
    assert np.all(dr[:, :, L["delta_z"]:L["delta_z"] + model.nparam - 1] == 0.0)
    xg = np.linspace(-1.9, 1.9, 41)
    j, vals = lp.banded_basis(xg, basis.knots)
    D = np.zeros((xg.shape[0], basis.nbases + 1))
    for m in range(4):
        D[np.arange(xg.shape[0]), j - 3 + m] = vals[:, m]
    DZ = D[:, :basis.nbases] @ basis.Z
    beta = dr[:, :, L["beta"][0]:L["beta"][0] + p].reshape(-1, p)
    f_hat = (DZ @ beta.T).mean(axis=1)
    truth = 0.8 * np.sin(1.5 * xg)
    rmse = np.sqrt(np.mean(((f_hat - f_hat.mean()) - (truth - truth.mean())) ** 2))
    assert rmse < 0.08, rmse
    sigma_hat = np.exp(dr[:, :, L["gamma0"]]).mean()
    assert abs(sigma_hat - 0.5) < 0.05, sigma_hat
    assert abs(dr[:, :, L["beta0"]].mean() - y.mean()) < 0.05
