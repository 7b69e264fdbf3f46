"""Callers of the hot path on new data against the oracle:
predictive z / log-density / CDF (`ptm_predict_*`), predictive quantiles
(`ptm_predict_quantile_*`) and the inverse transformation (`ptm_inverse_*`)."""

import numpy as np
import pytest
from scipy.special import ndtri

from oracle import ptm_oracle as o
from tests.helpers import build_case, rel_err

torch = pytest.importorskip("torch")


# ---------------------------------------------------------------------------------------
# oracle properties (CPU)
# ---------------------------------------------------------------------------------------
def test_oracle_inverse_round_trip():
    """tests/test_bspline/test_ptm.py:167-301: h^{-1}(h(x)) = x within 1e-5; and the inverse is
    exact for the forward function: h(h^{-1}(z)) = z to rounding."""
    kn = o.PTMKnots(-4.0, 4.0, 20)
    sp = o.PTMSpline(kn.knots)
    rng = np.random.default_rng(0)
    for delta in (np.zeros(20), rng.normal(size=20), 0.3 * rng.normal(size=20)):
        x = np.concatenate([np.linspace(-8, 8, 2001), rng.normal(scale=3, size=500),
                            [sp.min_eps, sp.max_eps, sp.min_knot, sp.max_knot]])
        z, _ = sp.dot_and_deriv_n_fullbatch(x, delta)
        r = sp.dot_inverse_exact(z, delta)
        assert np.max(np.abs(r - x)) < 1e-5
        z2, _ = sp.dot_and_deriv_n_fullbatch(r, delta)
        assert np.max(np.abs(z2 - z)) < 1e-13 * max(1.0, np.max(np.abs(z)))
        zs = np.sort(z)
        assert np.all(np.diff(sp.dot_inverse_exact(zs, delta)) >= 0)
    assert np.isnan(sp.dot_inverse_exact(np.array([np.nan]), np.zeros(20))[0])
    # identity transformation at delta = 0 (tests/test_predictor.py:14-22) -- up to the
    # 1000/999 lerp-weight quirk of the forward (approx.py:375-376): 0.1 % of a grid cell
    zz = np.linspace(-7, 7, 101)
    assert np.max(np.abs(sp.dot_inverse_exact(zz, np.zeros(20)) - zz)) < 1e-5


# ---------------------------------------------------------------------------------------
# CUDA path (GPU)
# ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_inverse_vs_oracle(dtype, tol):
    case = build_case(n=8, n_loc=1, n_scale=0, nbases=8, nparam=20, C=1, dtype=dtype)
    rng = np.random.default_rng(4)
    delta = np.stack([np.zeros(20), rng.normal(size=20), 0.3 * rng.normal(size=20)]).astype(dtype)
    sp = o.PTMSpline(o.PTMKnots(-4.0, 4.0, 20).knots)
    x = np.concatenate([np.linspace(-8, 8, 3001), rng.normal(scale=3, size=1000),
                        [sp.min_eps, sp.max_eps, sp.min_knot, sp.max_knot]])
    for c in range(3):
        z64, _ = sp.dot_and_deriv_n_fullbatch(x, delta[c].astype(np.float64))
        z = z64.astype(dtype)
        r = case.model.inverse(torch.from_numpy(delta[c]).cuda(), torch.from_numpy(z).cuda())
        r_ref = sp.dot_inverse_exact(z.astype(np.float64), delta[c].astype(np.float64))
        scale = np.max(np.abs(r_ref))
        if dtype == np.float64:
            # the forward is a saw-tooth of 0.1 % of a grid cell at every grid point
            # (lerp weight up to 1000/999, approx.py:375-376), so a z within rounding of a
            # sample H[i] has two exact pre-images 8e-6 apart: accept either
            err = np.abs(r.cpu().numpy() - r_ref)
            fwd, _ = sp.dot_and_deriv_n_fullbatch(r.cpu().numpy(), delta[c].astype(np.float64))
            twin = (err < 8.1e-6) & (np.abs(fwd - z) < 1e-13 * max(1.0, np.max(np.abs(z))))
            assert np.all((err < tol * scale) | twin) and twin.sum() - (err < tol * scale)[twin].sum() < 8
        else:  # float cells are decided on float samples: compare through the round trip
            assert np.max(np.abs(r.cpu().numpy() - r_ref)) < 1e-4
        # round trip through the CUDA forward
        z2, _ = case.model.transform(torch.from_numpy(delta[c]).cuda(), r)
        assert np.max(np.abs(z2.cpu().numpy() - z)) < 10 * tol * max(1.0, np.max(np.abs(z)))
        # the reference's own round-trip tolerance (test_ptm.py:167-301)
        assert np.max(np.abs(r.cpu().numpy() - x)) < (1e-5 if dtype == np.float64 else 1e-4)
    nan = case.model.inverse(torch.from_numpy(delta).cuda(), torch.tensor([float("nan")], dtype=r.dtype).cuda())
    assert torch.isnan(nan).all() and nan.shape == (3, 1)
    with pytest.raises(TypeError):
        case.model.inverse(torch.from_numpy(delta[0]).cuda(), torch.zeros((2, 2), dtype=r.dtype).cuda())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_summarise_dist_newdata_vs_oracle(dtype, tol):
    case = build_case(n=300, n_loc=2, n_scale=1, nbases=12, nparam=20, C=4, dtype=dtype, amp=0.3)
    rng = np.random.default_rng(9)
    m = 1537
    newdata = {f"x{t}": rng.uniform(-1.99, 1.99, size=m).astype(dtype) for t in range(3)}
    grid = rng.normal(scale=2.5, size=m).astype(dtype)
    grid[:4] = [9.0, -9.0, 4.4, -4.3]
    grid[7] = np.nan
    params = torch.from_numpy(case.params).cuda()
    out = case.model.summarise_dist(params, grid=grid, newdata=newdata)
    om = case.omodel
    st = case.state
    if dtype == np.float32:  # the oracle in double on the float-rounded inputs
        from tests import golden_io
        om = o.PTMModel(case.omodel.y.astype(np.float64),
                        [o.term_from_reparam(t.x.astype(np.float64), t.knots.astype(np.float64),
                                             t.Z.astype(np.float64), t.K.astype(np.float64), t.rank, t.predictor)
                         for t in case.omodel.terms], nparam=20)
        om.Zd = case.omodel.Zd.astype(np.float64)
        st = golden_io.state_from_params(om, case.params.astype(np.float64), case.layout)
    X = [newdata[f"x{t}"].astype(np.float64) for t in range(3)]
    for c in range(4):
        z, lp, cdf, _, _ = om.predict(st, c, grid.astype(np.float64), X)
        assert rel_err(out["z"][c].cpu().numpy(), z) < tol
        assert rel_err(out["log_prob"][c].cpu().numpy(), lp) < tol
        assert rel_err(out["cdf"][c].cpu().numpy(), cdf) < max(tol, 1e-14) * 10
        assert rel_err(out["prob"][c].cpu().numpy(), np.exp(lp)) < tol * 10
        assert np.isnan(out["log_prob"][c, 7].item()) and np.isnan(out["cdf"][c, 7].item())
    assert out["z"].shape == (4, m)
    # training data by default: the per-sample sums are the log-likelihood of log_prob()
    if dtype == np.float64:
        tr = case.model.summarise_dist(params, which=("log_prob",))
        assert set(tr) == {"log_prob"}
        pri = np.array([case.omodel.prior(case.state, c) for c in range(4)])
        full = case.model.log_prob(params).cpu().numpy()
        assert rel_err(tr["log_prob"].sum(dim=1).cpu().numpy() + pri, full) < 1e-12
    # Basis.bspline raises outside the interior knots (approx.py:197-206)
    bad = dict(newdata)
    bad["x1"] = bad["x1"].copy()
    bad["x1"][3] = 2.5
    with pytest.raises(ValueError):
        case.model.summarise_dist(params, grid=grid, newdata=bad)
    with pytest.raises(KeyError):
        case.model.summarise_dist(params, grid=grid, newdata={"x0": newdata["x0"]})


@pytest.mark.gpu
def test_predict_quantile_vs_oracle():
    case = build_case(n=300, n_loc=2, n_scale=1, nbases=12, nparam=20, C=5, amp=0.3)
    rng = np.random.default_rng(10)
    m = 777
    newdata = {f"x{t}": rng.uniform(-1.99, 1.99, size=m) for t in range(3)}
    u = rng.uniform(1e-6, 1 - 1e-6, size=m)
    u[:3] = [1e-12, 1 - 1e-12, 0.5]
    params = torch.from_numpy(case.params).cuda()
    yq = case.model.predict_quantile(params, u, newdata=newdata).cpu().numpy()
    X = [newdata[f"x{t}"] for t in range(3)]
    zq = ndtri(u)
    for c in range(5):
        _, _, _, mu, sd = case.omodel.predict(case.state, c, np.zeros(m), X)
        delta = case.state.delta_z[c] @ case.omodel.Zd.T
        ref = mu + sd * case.omodel.spline.dot_inverse_exact(zq, delta)
        assert rel_err(yq[c], ref) < 1e-11
    # cdf(quantile(u)) = u, sample by sample
    for c in range(2):
        cdf = case.model.summarise_dist(params[c:c + 1], grid=yq[c], newdata=newdata, which=("cdf",))["cdf"]
        assert np.max(np.abs(cdf.cpu().numpy()[0] - u)) < 1e-11
    # scalar probability broadcasts over the rows
    med = case.model.predict_quantile(params, 0.5, newdata=newdata)
    assert med.shape == (5, m)


@pytest.mark.gpu
@pytest.mark.parametrize("n_loc,n_scale,C,m", [(1, 0, 1, 5), (6, 3, 7, 300), (8, 5, 3, 1000), (2, 2, 37, 129)])
def test_predict_shapes_vs_oracle(n_loc, n_scale, C, m):
    """Register tiers of the observation-major sweep (T <= 4, 8, 12), the sample-major sweep
    beyond 12 terms, ragged sample groups and a ragged last observation block."""
    case = build_case(n=200, n_loc=n_loc, n_scale=n_scale, nbases=10, nparam=20, C=C, amp=0.3)
    T = n_loc + n_scale
    rng = np.random.default_rng(20 + T)
    newdata = {f"x{t}": rng.uniform(-1.99, 1.99, size=m) for t in range(T)}
    grid = rng.normal(scale=2.5, size=m)
    u = rng.uniform(0.01, 0.99, size=m)
    params = torch.from_numpy(case.params).cuda()
    out = case.model.summarise_dist(params, grid=grid, newdata=newdata)
    yq = case.model.predict_quantile(params, u, newdata=newdata).cpu().numpy()
    X = [newdata[f"x{t}"] for t in range(T)]
    zq = ndtri(u)
    for c in range(C):
        z, lp, cdf, mu, sd = case.omodel.predict(case.state, c, grid, X)
        assert rel_err(out["z"][c].cpu().numpy(), z) < 1e-12
        assert rel_err(out["log_prob"][c].cpu().numpy(), lp) < 1e-12
        assert rel_err(out["cdf"][c].cpu().numpy(), cdf) < 1e-12
        delta = case.state.delta_z[c] @ case.omodel.Zd.T
        ref = mu + sd * case.omodel.spline.dot_inverse_exact(zq, delta)
        assert rel_err(yq[c], ref) < 1e-11


@pytest.mark.gpu
def test_predictive_draws_are_quantiles_of_per_sample_uniforms():
    """init_dist(samples).sample() (TransformationDist._sample_n, dist.py:193-201): one uniform
    per (posterior sample, row), pushed through the quantile function.  The draw with a given
    u [S, m] equals the quantile sweep row by row, maps back to u through the predictive CDF,
    and a seeded generator reproduces it."""
    import torch

    case = build_case(n=300, n_loc=2, n_scale=1, nbases=10, C=6)
    model = case.model
    samples = torch.from_numpy(case.params).cuda()
    S, m = 6, 300
    gen = torch.Generator(device="cuda")
    gen.manual_seed(7)
    u = 1e-3 + (1 - 2e-3) * torch.rand((S, m), dtype=torch.float64, device="cuda", generator=gen)
    y = model.sample(samples, u=u)
    assert y.shape == (S, m) and torch.isfinite(y).all()
    for s in (0, 5):  # row s of the draw = quantile sweep at u[s] for sample s
        q = model.predict_quantile(samples[s:s + 1], u[s].cpu().numpy())
        assert rel_err(y[s].cpu().numpy(), q[0].cpu().numpy()) < 1e-13
    # CDF of the draw under its own sample recovers u (forward o inverse, up to the reference's
    # 1000 / 999 lerp-weight quirk of the forward function)
    for s in (1, 4):
        cdf = model.summarise_dist(samples[s:s + 1], grid=y[s].cpu().numpy(), which=("cdf",))["cdf"]
        assert np.max(np.abs(cdf[0].cpu().numpy() - u[s].cpu().numpy())) < 5e-6
    g1 = torch.Generator(device="cuda"); g1.manual_seed(11)
    g2 = torch.Generator(device="cuda"); g2.manual_seed(11)
    assert torch.equal(model.sample(samples, generator=g1), model.sample(samples, generator=g2))
