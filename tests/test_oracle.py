"""CPU tests of the oracle: the properties the reference's own tests pin (SURVEY.md
section 4 / 8c), the gradient against a literal JVP restatement, and regression
against the committed golden fixtures."""

import numpy as np
import pytest

from oracle import ptm_oracle as o
from tests import golden_io
from tests.helpers import build_case, oracle_eval, pack_grad, rel_err

knots10 = o.PTMKnots(-4.0, 4.0, nparam=10)
bs10 = o.PTMSpline(knots10.knots)
coef10 = np.random.default_rng(1).normal(size=10)


class TestDotAndDeriv:
    """tests/test_bspline/test_ptm.py:12-164 of the reference."""

    def test_scalar_x(self):
        fx, fxd = bs10.dot_and_deriv_n_fullbatch(np.array(1.0), coef10)
        assert np.shape(fx) == () and np.shape(fxd) == ()
        assert not np.isnan(fx) and not np.isnan(fxd)
        assert fxd > 0.0

    @pytest.mark.parametrize("m", [300, 10_000])
    def test_vector_x(self, m):
        x = np.linspace(-8.0, 8.0, m)
        fx, fxd = bs10.dot_and_deriv_n_fullbatch(x, coef10)
        assert fx.shape == x.shape and fxd.shape == x.shape
        assert not np.isnan(fx).any() and not np.isnan(fxd).any()
        assert np.all(fxd > 0.0)
        assert np.all(np.diff(fx) > 0.0)

    def test_fullbatch_rejects_batched(self):
        """tests/test_bspline/test_ptm.py:335-350."""
        rng = np.random.default_rng(1)
        with pytest.raises(TypeError):
            bs10.dot_and_deriv_n_fullbatch(rng.normal(size=(4, 200)), coef10)
        with pytest.raises(ValueError):
            bs10.dot_and_deriv_n_fullbatch(np.array(1.0), rng.normal(size=(3, 10)))
        with pytest.raises(ValueError):
            bs10.dot_and_deriv_n_fullbatch(rng.normal(size=200), rng.normal(size=(3, 10)))


def test_identity_at_zero():
    """tests/test_predictor.py:14-22: theta(0) = [knots[2], step, ...] => h = identity."""
    theta = bs10.compute_coef(np.zeros(10))
    assert theta[0] == pytest.approx(knots10.knots[2], abs=1e-6)
    assert np.allclose(theta[1:], knots10.step, atol=1e-6)
    assert theta.shape[-1] == 11
    x = np.linspace(-8, 8, 400)
    fx, fxd = bs10.dot_and_deriv_n_fullbatch(x, np.zeros(10))
    assert np.max(np.abs(fx - x)) < 5e-5  # the grid-lerp error of the reference
    assert np.max(np.abs(fxd - 1.0)) < 1e-12


def test_grid_step_quirk():
    """approx.py:375-376: spacing (b-a)/999 but step (b-a)/1000 => weight up to 1.001."""
    ap = bs10.bspline
    assert np.isclose(np.diff(ap.grid[1:-1]).mean(), 8.0 / 999)
    assert np.isclose(ap.step, 8.0 / 1000)
    x = np.nextafter(ap.grid[5], -np.inf)
    _, _, _, k = ap.cell(np.array([x]))
    assert 1.0 < k[0] < 1.0011


def test_derivative_tables_are_derivatives():
    """G' theta and G'' theta are the exact derivatives of the spline G theta (equidistant)."""
    ap = bs10.bspline
    theta = bs10.compute_coef(coef10)
    xs = np.linspace(-3.9, 3.9, 41)
    h = 1e-5
    B = lambda v: o.bspline_basis(v, ap.knots) @ ap.postmultiply_by @ theta  # noqa: E731
    num_d = (B(xs + h) - B(xs - h)) / (2 * h)
    num_d2 = (B(xs + h) - 2 * B(xs) + B(xs - h)) / h**2
    d = o.bspline_basis_deriv(xs, ap.knots) @ ap.postmultiply_by @ theta
    d2 = o.bspline_basis_deriv2(xs, ap.knots) @ ap.postmultiply_by @ theta
    assert np.max(np.abs(d - num_d)) < 1e-8
    assert np.max(np.abs(d2 - num_d2)) < 1e-4


class TestConstraints:
    """tests/test_constraint.py:16-67."""

    x = np.random.default_rng(1).uniform(size=15)
    knots = o.equidistant_knots(x, n_param=7)
    basis = o.bspline_basis(x, knots)

    @pytest.mark.parametrize("seed", range(5))
    def test_sumzero_coef(self, seed):
        Cbar = o.sumzero_coef(7)
        a = np.random.default_rng(seed).normal(size=Cbar.shape[-1])
        assert (Cbar @ a).sum() == pytest.approx(0.0, abs=1e-4)

    @pytest.mark.parametrize("seed", range(5))
    def test_sumzero_term(self, seed):
        Cbar = o.sumzero_term(self.basis)
        a = np.random.default_rng(seed).normal(size=Cbar.shape[-1])
        assert (self.basis @ Cbar @ a).sum() == pytest.approx(0.0, abs=1e-4)


def test_ptmcoef_latent_dim_and_penalty():
    """tests/test_coef.py:7-23: latent dim nparam-1 and penalty == I after diagonalisation."""
    Z = o.trafo_Z(10)
    assert Z.shape == (10, 9)
    pen = o.pspline_penalty(10, 1)
    pen = pen / np.linalg.norm(pen, ord=np.inf)
    assert np.allclose(Z.T @ pen @ Z, np.eye(9), atol=1e-8)


def test_ps_term_sums_to_zero():
    """tests/test_gam/test_var.py:359-369."""
    x = np.random.default_rng(0).uniform(-2, 2, 200)
    t = o.ps(x, 12)
    a = np.random.default_rng(1).normal(size=t.p)
    assert abs((t.design @ a).sum()) < 1e-8
    assert t.rank == t.p - 1
    assert np.allclose(t.design, o.basis_matrix(x, t.knots) @ t.Z)


def test_mvn_singular_matches_degenerate_up_to_constant():
    """tests/test_gam/test_dist.py:12-24: equal to a degenerate MVN log-density up to a
    constant (here: the generic pseudo-determinant form)."""
    pen = o.pspline_penalty(10, 2)
    x = np.random.default_rng(1).normal(size=(20, 10))
    lp = o.mvn_singular_logp(x, pen, 1.0, 8)
    ev = np.linalg.eigvalsh(pen)
    logpdet = np.sum(np.log(ev[ev > 1e-8]))
    ref = -0.5 * np.einsum("ni,ij,nj->n", x, pen, x) + 0.5 * logpdet - 0.5 * 8 * np.log(2 * np.pi)
    assert np.allclose(np.diff(ref - lp), 0.0, atol=1e-9)


def test_basis_raises_outside_knots():
    """bspline/approx.py:197-206."""
    knots = o.equidistant_knots(np.array([-2.0, 2.0]), 10)
    with pytest.raises(ValueError):
        o.basis_matrix(np.array([0.0, 2.5]), knots)


def test_gradient_matches_jvp_transcription():
    case = build_case(n=400, n_loc=2, n_scale=1, nbases=12, C=2, build=False)
    lp, g = case.omodel.logp_and_grad(case.state)
    rng = np.random.default_rng(7)
    st = case.state
    for c in range(2):
        d = o.ChainState(beta=[rng.normal(size=b.shape) for b in st.beta], tau=rng.normal(size=st.tau.shape),
                         delta_z=rng.normal(size=st.delta_z.shape), tau_delta=rng.normal(size=2),
                         beta0=rng.normal(size=2), gamma0=rng.normal(size=2))
        jv = case.omodel.jvp_logp(st, d, c)
        gv = sum((g["beta"][t][c] * d.beta[t][c]).sum() for t in range(3)) + (g["tau"][c] * d.tau[c]).sum()
        gv += (g["delta_z"][c] * d.delta_z[c]).sum() + g["tau_delta"][c] * d.tau_delta[c]
        gv += g["beta0"][c] * d.beta0[c] + g["gamma0"][c] * d.gamma0[c]
        assert abs(jv - gv) < 1e-10 * max(1.0, abs(jv))


def test_logp_matches_direct_dense_evaluation():
    """logp re-derived from scratch with the dense tables: sum of Normal logpdf + logdet + priors."""
    case = build_case(n=300, n_loc=1, n_scale=1, nbases=10, C=1, build=False)
    om, st = case.omodel, case.state
    lp = om.logp(st)[0]
    mu, ls = om.predictors(st, 0)
    r = (om.y - mu) / np.exp(ls)
    z, d = om.spline.dot_and_deriv_n_fullbatch(r, st.delta_z[0] @ om.Zd.T)
    ll = -0.5 * z**2 - 0.5 * np.log(2 * np.pi) + np.log(d) - ls
    assert abs(lp - (ll.sum() + om.prior(st, 0))) < 1e-9 * abs(lp)


GOLDEN_LOGP = ["n1_c1", "n7_c3", "n100_c1", "n4097_c3", "nan_in_y", "far_tails", "xtwx_nb20", "xtwx_nb30"]


@pytest.mark.parametrize("name", GOLDEN_LOGP)
def test_oracle_reproduces_golden(name):
    d = golden_io.load(f"logp_{name}.npz")
    om = golden_io.oracle_from_fixture(d)
    model = golden_io.model_from_fixture(d, device="cpu")
    layout, P = model.param_layout()
    st = golden_io.state_from_params(om, d["params"], layout)
    lp, g = om.logp_and_grad(st)

    class _C:
        pass

    c = _C()
    c.layout, c.P = layout, P
    assert rel_err(lp, d["logp"]) < 1e-13
    assert rel_err(pack_grad(c, g), d["grad"]) < 1e-13
    if name == "nan_in_y":
        assert np.isnan(d["logp"]).all()
    if name.startswith("xtwx"):
        x, Pm, L = om.loc_precision(st, 0)
        assert rel_err(x, d["xtwx"]) < 1e-13 and rel_err(L, d["chol"]) < 1e-12


def test_golden_basis_and_trafo():
    d = golden_io.load("basis_nb20.npz")
    assert np.array_equal(o.basis_matrix(d["x"], d["knots"]), d["dense"])
    t = golden_io.load("trafo_nparam20.npz")
    sp = o.PTMSpline(t["knots"])
    assert np.array_equal(sp.bspline.grid, t["grid"])
    assert np.allclose(sp.compute_coef(t["delta"]), t["theta"], rtol=1e-14)
    for c in range(3):
        z, dd = sp.dot_and_deriv_n_fullbatch(t["r"], t["delta"][c])
        assert np.allclose(z, t["z"][c], rtol=1e-14, atol=1e-14)
    assert np.array_equal(sp.region_code(t["r"]).astype(np.int32), t["code"])


def test_noncentered_form_is_the_centred_density_up_to_its_prior():
    """gam/var.py:176-197, var.py:100-107: with coef = scale * latent the likelihood is the
    centred one at coef, the prior is N(0, inv(K)) / N(0, 1) on the latent vector, and the
    gradient follows by the chain rule (checked against central differences too)."""
    y, X = o.synth_data(400, 1, 1, seed=2)
    terms = [o.ps(X[0], 10, predictor="loc"), o.ps(X[1], 10, predictor="scale")]
    om = o.PTMModel(y, terms, nparam=12)
    st = o.synth_state(terms, 12, 2, seed=3, amp=0.2)
    st.tau[:] = np.array([[0.7, 1.3], [1.1, 0.8]])
    st.tau_delta[:] = np.array([0.6, 0.9])
    lp_c, _ = om.logp_and_grad(st, add_priors=False)
    # the same coefficients expressed through latent vectors
    lat = o.ChainState(beta=[st.beta[t] / st.tau[:, t:t + 1] for t in range(2)], tau=st.tau.copy(),
                       delta_z=st.delta_z / st.tau_delta[:, None], tau_delta=st.tau_delta.copy(),
                       beta0=st.beta0.copy(), gamma0=st.gamma0.copy())
    for t in terms:
        t.noncentered = True
    om.trafo_noncentered = True
    lp_n, g = om.logp_and_grad(lat, add_priors=False)
    assert np.allclose(lp_n, lp_c, rtol=1e-13)
    lp_full = om.logp(lat)
    prior = lp_full - lp_n
    for c in range(2):
        ref = sum(o.mvn_singular_logp(lat.beta[t][c], terms[t].K, 1.0, terms[t].rank) for t in range(2))
        ref += float(o.normal_iid_logp(lat.delta_z[c][None, :], np.ones(1))[0])
        assert abs(prior[c] - ref) < 1e-9
    # chain rule against the CENTRED gradient (which the reference-code fixture pins): the
    # declared derivatives of the custom_jvp are not finite-difference slopes, so the check is
    # algebraic
    for t in terms:
        t.noncentered = False
    om.trafo_noncentered = False
    _, gc = om.logp_and_grad(st, add_priors=False)
    for t in range(2):
        assert np.allclose(g["beta"][t], st.tau[:, t:t + 1] * gc["beta"][t], rtol=1e-12, atol=1e-12)
        assert np.allclose(g["tau"][:, t], np.sum(lat.beta[t] * gc["beta"][t], axis=1), rtol=1e-12, atol=1e-12)
    assert np.allclose(g["delta_z"], st.tau_delta[:, None] * gc["delta_z"], rtol=1e-12, atol=1e-12)
    assert np.allclose(g["tau_delta"], np.sum(lat.delta_z * gc["delta_z"], axis=1), rtol=1e-12, atol=1e-12)
