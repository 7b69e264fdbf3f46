#!/usr/bin/env python3
"""Golden fixtures produced by the REFERENCE'S OWN SOURCE FILES (tests/golden/ref_*.npz).

jax / tfp / liesel are not installed in this image, so the reference cannot be imported as
a package.  This script instead executes the reference's hot-path files unmodified, loaded
by path from /root/reference/liesel_ptm, over the NumPy stand-ins in tests/golden/jaxshim
(see its README): bspline/approx.py, bspline/ptm.py, bspline/util.py, dist.py, gam/dist.py,
constraint.py, penalty.py, iwls_proposals.py, predictor.py (static intercept formulas).  The package `__init__`s are bypassed (they import the liesel
graph machinery); everything evaluated below is the reference's code path for
`LocScaleTransformationDist(..., batched=False).log_prob` and its building blocks.

Run in the build container only (needs /root/reference):

    python tests/golden/make_reference_golden.py

The fixtures travel to the GPU box; the tests never import the shim or the reference.
"""

import importlib
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PTM_REFERENCE_ROOT", "/root/reference")
PKG = os.path.join(REF, "liesel_ptm")


def load_reference():
    """Reference modules by path, with the NumPy stand-ins first on sys.path."""
    sys.path.insert(0, os.path.join(HERE, "jaxshim"))
    import jax  # noqa: F401  (the stand-in)
    import tensorflow_probability  # noqa: F401

    assert "jaxshim" in jax.__file__

    def fake_pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        m.__package__ = name
        sys.modules[name] = m
        return m

    fake_pkg("liesel_ptm", PKG)
    bs = fake_pkg("liesel_ptm.bspline", os.path.join(PKG, "bspline"))
    fake_pkg("liesel_ptm.util", os.path.join(PKG, "util"))
    fake_pkg("liesel_ptm.gam", os.path.join(PKG, "gam"))

    mods = types.SimpleNamespace()
    mods.approx = importlib.import_module("liesel_ptm.bspline.approx")
    mods.util = importlib.import_module("liesel_ptm.bspline.util")
    mods.ptm = importlib.import_module("liesel_ptm.bspline.ptm")
    bs.PTMSpline = mods.ptm.PTMSpline
    bs.PTMKnots = mods.ptm.PTMKnots
    bs.OnionSpline = type("OnionSpline", (), {})  # off-path variant, name only
    mods.dist = importlib.import_module("liesel_ptm.dist")
    mods.gamdist = importlib.import_module("liesel_ptm.gam.dist")
    mods.constraint = importlib.import_module("liesel_ptm.constraint")
    mods.penalty = importlib.import_module("liesel_ptm.penalty")
    # iwls_proposals.py only needs the NAMES FlatLogProb / PTMCoef at import time (they
    # belong to the observed-information providers, which are not on this path)
    for name, attr in (("liesel_ptm.logprob", "FlatLogProb"), ("liesel_ptm.var", "PTMCoef")):
        stub = types.ModuleType(name)
        setattr(stub, attr, type(attr, (), {}))
        sys.modules[name] = stub
    mods.iwls = importlib.import_module("liesel_ptm.iwls_proposals")
    del sys.modules["liesel_ptm.logprob"], sys.modules["liesel_ptm.var"]
    # predictor.py: only the static intercept formulas are called (the classes derive from
    # inert liesel.model placeholders)
    mods.predictor = importlib.import_module("liesel_ptm.predictor")
    for m in vars(mods).values():
        assert m.__file__.startswith(REF), m.__file__
    return mods


def A(x):
    return np.asarray(x, dtype=np.float64)


def ref_basis(m, out):
    """K1: equidistant_knots + Basis.bspline / bspline_basis on and around every knot."""
    nb = 20
    rng = np.random.default_rng(11)
    xdata = rng.uniform(-2, 2, size=400)
    knots = m.approx.equidistant_knots(xdata, nb, order=3, eps=0.01)
    kn = A(knots)
    xs = []
    for j in range(3, nb + 1):
        k = kn[j]
        for v in (np.nextafter(k, -np.inf), k, np.nextafter(k, np.inf)):
            if kn[3] <= v <= kn[nb]:
                xs.append(v)
    xs += list(xdata[:120])
    x = np.array(xs)
    dense = A(m.approx.Basis.bspline(x, knots, 3))
    x_out = np.concatenate([x[:40], [kn[3] - 0.3, kn[nb] + 0.2, kn[0] - 1.0]])
    masked = A(m.approx.bspline_basis(x_out, knots, 3))
    d1 = A(m.approx.bspline_basis_deriv(x[:60], knots, 3))
    d2 = A(m.approx.bspline_basis_deriv2(x[:60], knots, 3))
    try:
        m.approx.Basis.bspline(np.array([kn[3] - 0.1]), knots, 3)
        raised = False
    except ValueError:
        raised = True
    np.savez(os.path.join(out, "ref_basis_nb20.npz"), xdata=xdata, knots=kn, x=x, dense=dense,
             x_out=x_out, masked=masked, deriv=d1, deriv2=d2, raised=np.array(raised))


def ref_trafo(m, out):
    """PTMKnots, PTMCoef map, BSplineApprox tables, PTMSpline over all five regions,
    and the declared tangent rule of the custom_jvp."""
    nparam = 20
    kn = m.ptm.PTMKnots(-4.0, 4.0, nparam)
    sp = m.ptm.PTMSpline(kn.knots, eps=0.1)
    ap = sp.bspline
    g = A(ap.grid)
    rng = np.random.default_rng(5)
    r = np.concatenate([
        g[::7], np.nextafter(g[::7], -np.inf), np.nextafter(g[::7], np.inf), g[-4:], g[:4],
        [float(sp.min_eps), float(sp.max_eps),
         np.nextafter(float(sp.min_eps), -np.inf), np.nextafter(float(sp.max_eps), np.inf)],
        np.linspace(-9, 9, 301), rng.normal(scale=2.5, size=300),
    ])
    delta = np.stack([np.zeros(nparam), rng.normal(size=nparam), 0.3 * rng.normal(size=nparam)])
    theta = np.stack([A(sp.compute_coef(d)) for d in delta])
    z = np.zeros((3, r.size))
    d = np.zeros_like(z)
    for c in range(3):
        zz, dd = sp.dot_and_deriv_n_fullbatch(r, delta[c])
        z[c], d[c] = A(zz), A(dd)

    # declared derivatives (approx.py:498-520), core region only: unit tangents
    fn = ap._dot_and_deriv_fn
    assert fn.jvp is not None
    rc = r[(r >= float(ap.min_knot)) & (r <= float(ap.max_knot))][:200]
    J = theta.shape[1]
    h2 = np.zeros((3, rc.size))          # d h'/dr as declared
    dz_dr = np.zeros((3, rc.size))
    tb = np.zeros((3, rc.size, J))       # d z / d theta
    tbd = np.zeros((3, rc.size, J))      # d h'/ d theta
    for c in range(3):
        for i, x in enumerate(rc):
            (_, _), (t0, t1) = fn.jvp((x, theta[c]), (1.0, np.zeros(J)))
            dz_dr[c, i], h2[c, i] = float(t0), float(t1)
        for j in range(J):
            e = np.zeros(J)
            e[j] = 1.0
            for i, x in enumerate(rc):
                (_, _), (t0, t1) = fn.jvp((x, theta[c]), (0.0, e))
                tb[c, i, j], tbd[c, i, j] = float(t0), float(t1)
    cell = (np.searchsorted(g, r, side="right") - 1).astype(np.int64)
    np.savez(os.path.join(out, "ref_trafo_nparam20.npz"), knots=A(kn.knots), knot_step=A(kn.step),
             grid=g, grid_step=A(ap.step), basis_grid=A(ap.basis_grid),
             basis_deriv_grid=A(ap.basis_deriv_grid), basis_deriv2_grid=A(ap.basis_deriv2_grid),
             min_eps=A(sp.min_eps), max_eps=A(sp.max_eps), r=r, delta=delta, theta=theta,
             z=z, d=d, cell=cell, rc=rc, dz_dr=dz_dr, dd_dr=h2, dz_dtheta=tb, dd_dtheta=tbd)


def ref_logp(m, out):
    """Per-observation log-density through LocScaleTransformationDist(batched=False) with
    loc/scale built from the reference's basis, constraint and penalty code; priors through
    MultivariateNormalSingular; central-difference gradient of the reference-run logp."""
    import jax.numpy as jnp

    rng = np.random.default_rng(21)
    n, nb, nparam, C = 257, 12, 20, 2
    x_loc = rng.uniform(-2, 2, size=n)
    x_scale = rng.uniform(-2, 2, size=n)
    y = np.sin(x_loc) + np.exp(0.2 * np.cos(x_scale)) * np.where(
        rng.uniform(size=n) < 0.5, rng.normal(2, 1, size=n), rng.normal(-2, 0.5, size=n))
    y = (y - y.mean()) / y.std() * 1.6
    y[:6] = [9.5, -9.0, 4.3, -4.2, 5.1, -5.3]      # tails and transitions
    ynan = y.copy()
    ynan[17] = np.nan

    terms = []
    for x in (x_loc, x_scale):
        knots = m.approx.equidistant_knots(x, nb, order=3, eps=0.01)
        basis = m.approx.Basis.bspline(x, knots, 3)
        # ps() (gam/var.py:925-1022): pspline penalty, reparam_scale_penalty,
        # reparam_diagonalize_penalty, reparam_constraint("sumzero_term")
        K = m.penalty.Penalty.pspline(nparam=nb, random_walk_order=2)
        K = K / jnp.linalg.norm(K, ord=jnp.inf)
        Zd = m.constraint.mixed_model(K)
        Kd = Zd.T @ K @ Zd
        Zc = m.constraint.LinearConstraintEVD.sumzero_term(basis @ Zd)
        Kc = Zc.T @ Kd @ Zc
        Z = Zd @ Zc
        terms.append(dict(x=x, knots=A(knots), basis=A(basis), Z=A(Z), K=A(Kc),
                          rank=int(np.linalg.matrix_rank(A(Kc)))))

    kn = m.ptm.PTMKnots(-4.0, 4.0, nparam)
    sp = m.ptm.PTMSpline(kn.knots, eps=0.1)
    # var.py:66-74,202-212: latent trafo coefficients, delta = Z_delta delta_z
    Kt = m.penalty.Penalty.pspline(nparam=nparam, random_walk_order=1)
    Kt = Kt / jnp.linalg.norm(Kt, ord=jnp.inf)
    Z1 = m.constraint.mixed_model(Kt, rank=jnp.linalg.matrix_rank(Kt))[:, :-1]
    Z2 = m.constraint.mixed_model(jnp.eye(nparam - 1), rank=nparam - 1)
    Z_delta = A(Z1 @ Z2)

    P = 2 * (nb - 1) + (nparam - 1)
    p = nb - 1

    def unpack(v):
        return v[:p], v[p:2 * p], v[2 * p:]

    tau = np.array([[0.8, 1.3], [1.0, 0.6]])     # [chain][term]
    tau_d = np.array([0.5, 0.9])

    def loglik_i(v, yy):
        b_loc, b_scale, dz = unpack(v)
        mu = terms[0]["basis"] @ (terms[0]["Z"] @ b_loc)
        sigma = np.exp(terms[1]["basis"] @ (terms[1]["Z"] @ b_scale))
        delta = Z_delta @ dz
        dist = m.dist.LocScaleTransformationDist(
            coef=jnp.asarray(delta), loc=mu, scale=sigma, bspline=sp, batched=False)
        return A(dist.log_prob(jnp.asarray(yy)))

    def prior(v, c):
        b_loc, b_scale, dz = unpack(v)
        out_ = 0.0
        for t, b in enumerate((b_loc, b_scale)):
            d = m.gamdist.MultivariateNormalSingular(
                loc=0.0, scale=tau[c, t], penalty=terms[t]["K"], penalty_rank=terms[t]["rank"])
            out_ += float(d.log_prob(b))
        from tensorflow_probability.substrates.jax import distributions as tfd
        out_ += float(np.sum(tfd.Normal(loc=0.0, scale=tau_d[c]).log_prob(dz)))
        return out_

    def logp(v, c, yy):
        return float(np.sum(loglik_i(v, yy))) + prior(v, c)

    def grad_forward(v, c, yy):
        """Forward mode through the reference's code, one direction per parameter; the
        custom_jvp functions contribute their declared rule (approx.py:498-520)."""
        from jax._dual import Dual
        from tensorflow_probability.substrates.jax import distributions as tfd

        g = np.zeros(P)
        for k in range(P):
            e = np.zeros(P)
            e[k] = 1.0
            dv = Dual(v, e)
            b_loc, b_scale, dz = dv[:p], dv[p:2 * p], dv[2 * p:]
            mu = terms[0]["basis"] @ (terms[0]["Z"] @ b_loc)
            sigma = np.exp(terms[1]["basis"] @ (terms[1]["Z"] @ b_scale))
            delta = Z_delta @ dz
            dist = m.dist.LocScaleTransformationDist(
                coef=delta, loc=mu, scale=sigma, bspline=sp, batched=False)
            tot = np.sum(dist.log_prob(jnp.asarray(yy)))
            for t, b in enumerate((b_loc, b_scale)):
                d = m.gamdist.MultivariateNormalSingular(
                    loc=0.0, scale=tau[c, t], penalty=terms[t]["K"],
                    penalty_rank=terms[t]["rank"])
                tot = tot + d.log_prob(b)
            tot = tot + np.sum(tfd.Normal(loc=0.0, scale=tau_d[c]).log_prob(dz))
            if k == 0:
                assert abs(float(tot.p) - logp(v, c, yy)) <= 1e-9 * abs(float(tot.p))
            g[k] = float(tot.t)
        return g

    n_h = 96  # observations of the Hessian fixture (includes the tail / transition rows 0..5)

    def total_dual(dv, c, yy, rows):
        """The model log-density (likelihood of `rows` + priors) on a Dual parameter vector."""
        from tensorflow_probability.substrates.jax import distributions as tfd

        b_loc, b_scale, dz = dv[:p], dv[p:2 * p], dv[2 * p:]
        mu = terms[0]["basis"][rows] @ (terms[0]["Z"] @ b_loc)
        sigma = np.exp(terms[1]["basis"][rows] @ (terms[1]["Z"] @ b_scale))
        delta = Z_delta @ dz
        dist = m.dist.LocScaleTransformationDist(
            coef=delta, loc=mu, scale=sigma, bspline=sp, batched=False)
        tot = np.sum(dist.log_prob(jnp.asarray(yy[rows])))
        for t, b in enumerate((b_loc, b_scale)):
            d = m.gamdist.MultivariateNormalSingular(
                loc=0.0, scale=tau[c, t], penalty=terms[t]["K"], penalty_rank=terms[t]["rank"])
            tot = tot + d.log_prob(b)
        return tot + np.sum(tfd.Normal(loc=0.0, scale=tau_d[c]).log_prob(dz))

    def hessian_block(v, c, yy, lo, hi, full=False):
        """FlatLogProb.hessian for ONE coefficient block v[lo:hi] (logprob.py:104-113:
        `jacfwd(grad(log_prob))`): entry [i, j] = d/dv_j of (dlogp/dv_i).  grad is the inner
        derivative (highest Dual level, declared custom_jvp rule), jacfwd the outer one, which
        sees the rule body as ordinary code -- see jaxshim/jax/_dual.py.  The upper triangle
        is computed and mirrored unless `full` (one block is done in full to record that the
        jacfwd(grad) matrix is symmetric here)."""
        from jax._dual import Dual

        d = hi - lo
        H = np.zeros((d, d))
        rows = slice(0, n_h)
        for i in range(d):
            ei = np.zeros(P)
            ei[lo + i] = 1.0
            for j in range(d if full else i + 1):
                ej = np.zeros(P)
                ej[lo + j] = 1.0
                dv = Dual(Dual(v, ej, 1), ei, 2)
                tot = total_dual(dv, c, yy, rows)
                H[i, j] = float(tot.t.t)
                if not full:
                    H[j, i] = H[i, j]
        return H

    params = np.zeros((C, P))
    ll_i = np.zeros((C, n))
    ll_i_nan = np.zeros((C, n))
    lp = np.zeros(C)
    pri = np.zeros(C)
    grad_fd = np.zeros((C, P))
    grad = np.zeros((C, P))
    for c in range(C):
        v = np.concatenate([0.3 * rng.normal(size=p), 0.15 * rng.normal(size=p),
                            0.4 * rng.normal(size=nparam - 1)])
        params[c] = v
        ll_i[c] = loglik_i(v, y)
        ll_i_nan[c] = loglik_i(v, ynan)
        pri[c] = prior(v, c)
        lp[c] = logp(v, c, y)
        grad[c] = grad_forward(v, c, y)
        # Richardson-extrapolated central differences (4th order), step 1e-3:
        # smooth in v away from region/grid-cell switches; error ~1e-9 relative.
        for k in range(P):
            def f(h):
                e = np.zeros(P)
                e[k] = h
                return (logp(v + e, c, y) - logp(v - e, c, y)) / (2 * h)
            h = 1e-3
            grad_fd[c, k] = (4.0 * f(h / 2) - f(h)) / 3.0
    # Observed information (-Hessian) of the three coefficient blocks of chain 0 over the
    # first n_h observations + priors: what CholInfo / PTMCholInfoFixed / IWLSKernelFixed
    # .init_state / kernel.IWLSKernel._chol_info factorise (iwls_proposals.py:168-184,
    # 227-245, 279-298; iwls_fixed.py:118-143; kernel.py:240-258)
    hess_loc = hessian_block(params[0], 0, y, 0, p, full=True)
    hess_scale = hessian_block(params[0], 0, y, p, 2 * p)
    hess_dz = hessian_block(params[0], 0, y, 2 * p, P)

    if os.environ.get("PTM_GOLDEN_CHECK_HESS"):
        # self-check (off by default): the nested-Dual Hessian against central differences of
        # the first-order forward-mode gradient on the same rows (the gradient is piecewise
        # smooth: agreement ~1e-7 unless a row crosses a grid cell within the step)
        from jax._dual import Dual

        def grad_i(v, i):
            e = np.zeros(P)
            e[i] = 1.0
            return float(total_dual(Dual(v, e), 0, y, slice(0, n_h)).t)

        hcd = 1e-6
        for (blk, lo) in ((hess_loc, 0), (hess_scale, p), (hess_dz, 2 * p)):
            for (i, j) in ((0, 0), (1, 3), (4, 2), (7, 7)):
                ej = np.zeros(P)
                ej[lo + j] = hcd
                fd = (grad_i(params[0] + ej, lo + i) - grad_i(params[0] - ej, lo + i)) / (2 * hcd)
                print(f"hess check block@{lo} [{i},{j}]: nested {blk[i, j]:.10e}  fd {fd:.10e}")

    # IWLS information through GaussianLocCholInfo / GaussianScaleCholInfo
    # (iwls_proposals.py:55-89, 118-144) with a dict-backed model interface
    class DictModel:
        @staticmethod
        def extract_position(keys, state):
            return {k: state[k] for k in keys}

    prec = np.zeros((C, 2, p, p))
    chol = np.zeros((C, 2, p, p))
    for c in range(C):
        b_loc, b_scale, _ = unpack(params[c])
        sigma = np.exp(terms[1]["basis"] @ (terms[1]["Z"] @ b_scale))
        state = {"B_loc": jnp.asarray(terms[0]["basis"] @ terms[0]["Z"]),
                 "B_scale": jnp.asarray(terms[1]["basis"] @ terms[1]["Z"]),
                 "tau_loc": jnp.asarray(tau[c, 0]), "tau_scale": jnp.asarray(tau[c, 1]),
                 "sigma": jnp.asarray(sigma)}
        loc_info = m.iwls.GaussianLocCholInfo(
            basis_name="B_loc", smooth_name="f", smooth_scale_name="tau_loc", scale_name="sigma",
            penalty=jnp.asarray(terms[0]["K"]), model=DictModel, model_state=state)
        scale_info = m.iwls.GaussianScaleCholInfo(
            basis_name="B_scale", smooth_name="g", smooth_scale_name="tau_scale",
            penalty=jnp.asarray(terms[1]["K"]), model=DictModel, model_state=state)
        assert loc_info.n == n and scale_info.n == n
        prec[c, 0], chol[c, 0] = A(loc_info.precision(state)), A(loc_info.chol_info(state))
        prec[c, 1], chol[c, 1] = A(scale_info.precision(state)), A(scale_info.chol_info(state))

    # intercept pseudo-parameters (predictor.py:80-100, 126-130): loc_model / log_scale_model
    # are the predictors without their intercepts; the scale intercept sees the full loc
    icpt = np.zeros((C, 2))
    icpt_seq = np.zeros(C)  # gamma_0 when LogScaleIntercept sees the UPDATED location intercept
    for c in range(C):
        b_loc, b_scale, _ = unpack(params[c])
        loc_model = terms[0]["basis"] @ (terms[0]["Z"] @ b_loc)
        log_scale_model = terms[1]["basis"] @ (terms[1]["Z"] @ b_scale)
        icpt[c, 0] = float(m.predictor.LocationIntercept.compute_intercept(y, loc_model, log_scale_model))
        icpt[c, 1] = float(m.predictor.LogScaleIntercept.compute_intercept(y, loc_model, log_scale_model))
        # predictor.py:604-609: the scale intercept is wired to the full `loc` predictor, i.e. in
        # a sampler sweep that updates beta_0 first it sees loc = beta_0(new) + loc_model
        icpt_seq[c] = float(m.predictor.LogScaleIntercept.compute_intercept(
            y, icpt[c, 0] + loc_model, log_scale_model))

    np.savez(os.path.join(out, "ref_logp_n257.npz"), y=y, ynan=ynan, precision=prec, chol=chol,
             intercepts=icpt, intercepts_seq=icpt_seq,
             hess_n=n_h, hess_loc=hess_loc, hess_scale=hess_scale, hess_dz=hess_dz,
             x_loc=x_loc, x_scale=x_scale,
             knots_loc=terms[0]["knots"], knots_scale=terms[1]["knots"],
             Z_loc=terms[0]["Z"], Z_scale=terms[1]["Z"], K_loc=terms[0]["K"], K_scale=terms[1]["K"],
             rank_loc=terms[0]["rank"], rank_scale=terms[1]["rank"],
             basis_loc=terms[0]["basis"], basis_scale=terms[1]["basis"],
             Z_delta=Z_delta, trafo_knots=A(kn.knots),
             # the same data under the key names tests/golden_io.py reads
             nparam=nparam, n_terms=2, x0=x_loc, x1=x_scale, knots0=terms[0]["knots"],
             knots1=terms[1]["knots"], Z0=terms[0]["Z"], Z1=terms[1]["Z"], K0=terms[0]["K"],
             K1=terms[1]["K"], rank0=terms[0]["rank"], rank1=terms[1]["rank"], pred0=0, pred1=1,
             trafo_Z=Z_delta, tau=tau, tau_delta=tau_d, params=params,
             loglik_i=ll_i, loglik_i_nan=ll_i_nan, prior=pri, logp=lp, grad=grad, grad_fd=grad_fd)


def ref_variants(m, out):
    """PTMDist(bspline="onion" | "identity") (model/model.py:98-131) on the data and chain states
    of ref_logp_n257.npz: OnionKnots / get_onion_fn / OnionSpline (bspline/onion.py, run
    unmodified) through LocScaleTransformationDist, and GaussianPseudoTransformationDist
    (dist.py:766-850)."""
    import jax.numpy as jnp
    from jax._dual import Dual
    from tensorflow_probability.substrates.jax import distributions as tfd

    onion = importlib.import_module("liesel_ptm.bspline.onion")
    assert onion.__file__.startswith(REF), onion.__file__
    d0 = np.load(os.path.join(out, "ref_logp_n257.npz"))
    y, params, tau, tau_d, Z_delta = d0["y"], d0["params"], d0["tau"], d0["tau_delta"], d0["trafo_Z"]
    nparam = int(d0["nparam"])
    basis = [d0["basis_loc"], d0["basis_scale"]]
    Zs = [d0["Z0"], d0["Z1"]]
    Ks = [d0["K0"], d0["K1"]]
    ranks = [int(d0["rank0"]), int(d0["rank1"])]
    p = Zs[0].shape[1]
    C, P = params.shape
    n_h = int(d0["hess_n"])

    kn = onion.OnionKnots(-4.0, 4.0, nparam)
    sp = onion.OnionSpline(kn.knots)
    rng = np.random.default_rng(8)
    g = A(sp.bspline.grid)
    r = np.concatenate([g[::9], np.nextafter(g[::9], -np.inf), np.nextafter(g[::9], np.inf), g[-3:], g[:3],
                        np.linspace(-9, 9, 201), rng.normal(scale=2.5, size=200)])
    delta = np.stack([np.zeros(nparam), rng.normal(size=nparam), 0.3 * rng.normal(size=nparam)])
    theta = np.stack([A(sp.compute_coef(dd)) for dd in delta])
    z = np.zeros((3, r.size))
    dz = np.zeros_like(z)
    for c in range(3):
        zz, dd = sp.dot_and_deriv_n_fullbatch(r, delta[c])
        z[c], dz[c] = A(zz), A(dd)

    def make_dist(kind, delta_, mu, sigma):
        if kind == "onion":
            return m.dist.LocScaleTransformationDist(coef=delta_, loc=mu, scale=sigma, bspline=sp, batched=False)
        return m.dist.GaussianPseudoTransformationDist(coef=delta_, loc=mu, scale=sigma, batched=False)

    def total(kind, dv, c, rows, per_obs=False):
        b_loc, b_scale, dzv = dv[:p], dv[p:2 * p], dv[2 * p:]
        mu = basis[0][rows] @ (Zs[0] @ b_loc)
        sigma = np.exp(basis[1][rows] @ (Zs[1] @ b_scale))
        ll = make_dist(kind, Z_delta @ dzv, mu, sigma).log_prob(jnp.asarray(y[rows]))
        if per_obs:
            return ll
        tot = np.sum(ll)
        for t, b in enumerate((b_loc, b_scale)):
            dd = m.gamdist.MultivariateNormalSingular(loc=0.0, scale=tau[c, t], penalty=Ks[t], penalty_rank=ranks[t])
            tot = tot + dd.log_prob(b)
        if kind == "onion":  # the Gaussian model has no shape parameters (new_gaussian adds no trafo term)
            tot = tot + np.sum(tfd.Normal(loc=0.0, scale=tau_d[c]).log_prob(dzv))
        return tot

    res = {}
    allrows = slice(None)
    for kind in ("onion", "identity"):
        ll_i = np.zeros((C, y.shape[0]))
        lp = np.zeros(C)
        grad = np.zeros((C, P))
        for c in range(C):
            ll_i[c] = A(total(kind, params[c], c, allrows, per_obs=True))
            lp[c] = float(total(kind, params[c], c, allrows))
            for k in range(P):
                e = np.zeros(P)
                e[k] = 1.0
                t_ = total(kind, Dual(params[c], e), c, allrows)
                grad[c, k] = float(t_.t) if isinstance(t_, Dual) else 0.0
        res[kind] = (ll_i, lp, grad)

    def hessian_block(kind, v, c, lo, hi):
        dd = hi - lo
        H = np.zeros((dd, dd))
        rows = slice(0, n_h)
        for i in range(dd):
            ei = np.zeros(P)
            ei[lo + i] = 1.0
            for j in range(i + 1):
                ej = np.zeros(P)
                ej[lo + j] = 1.0
                tot = total(kind, Dual(Dual(v, ej, 1), ei, 2), c, rows)
                H[i, j] = H[j, i] = float(tot.t.t)
        return H

    hess = {"onion_loc": hessian_block("onion", params[0], 0, 0, p),
            "onion_scale": hessian_block("onion", params[0], 0, p, 2 * p),
            "onion_dz": hessian_block("onion", params[0], 0, 2 * p, P),
            "identity_loc": hessian_block("identity", params[0], 0, 0, p),
            "identity_scale": hessian_block("identity", params[0], 0, p, 2 * p)}
    np.savez(os.path.join(out, "ref_variants_n257.npz"),
             onion_knots=A(kn.knots), onion_step=A(kn.step), onion_grid=g, r=r, delta=delta, onion_theta=theta,
             onion_z=z, onion_d=dz, onion_min_knot=A(sp.min_knot), onion_max_knot=A(sp.max_knot),
             onion_loglik_i=res["onion"][0], onion_logp=res["onion"][1], onion_grad=res["onion"][2],
             identity_loglik_i=res["identity"][0], identity_logp=res["identity"][1],
             identity_grad=res["identity"][2], hess_n=n_h,
             **{f"hess_{k}": v for k, v in hess.items()})


def main():
    warnings.filterwarnings("ignore")
    np.seterr(all="ignore")
    out = os.environ.get("PTM_GOLDEN_OUT", HERE)  # tests regenerate into a scratch directory
    m = load_reference()
    ref_basis(m, out)
    ref_trafo(m, out)
    ref_logp(m, out)
    ref_variants(m, out)
    print("reference fixtures written to", out)


if __name__ == "__main__":
    main()
