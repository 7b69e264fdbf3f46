"""Parity against fixtures produced by the REFERENCE'S OWN SOURCE FILES.

tests/golden/ref_*.npz come from tests/golden/make_reference_golden.py, which executes
bspline/approx.py, bspline/ptm.py, bspline/util.py, dist.py, gam/dist.py, constraint.py,
penalty.py and iwls_proposals.py of /root/reference unmodified over NumPy stand-ins for
jax / tfp (tests/golden/jaxshim).  The gradient fixture is forward mode through that code
with the reference's declared custom_jvp rules (approx.py:498-520).

CPU part: the oracle (and the host mirror's setup code) against the fixtures.
GPU part: the CUDA path against the same fixtures, through the C ABI.
"""

import os

import numpy as np
import pytest

import liesel_ptm_b200 as lp
from oracle import ptm_oracle as o
from tests import golden_io
from tests.helpers import pack_grad, rel_err


def _state(d, C=2):
    p = d["Z0"].shape[1]
    P = d["params"]
    return o.ChainState(beta=[P[:, :p], P[:, p:2 * p]], tau=d["tau"], delta_z=P[:, 2 * p:],
                        tau_delta=d["tau_delta"], beta0=np.zeros(C), gamma0=np.zeros(C))


# ---------------------------------------------------------------------------------------
# oracle and host mirror vs the reference's code (CPU)
# ---------------------------------------------------------------------------------------
def test_oracle_basis_matches_reference_code_bit_for_bit():
    """equidistant_knots, Basis.bspline, bspline_basis(+deriv, deriv2): approx.py:11-276."""
    d = golden_io.load("ref_basis_nb20.npz")
    assert np.array_equal(o.equidistant_knots(d["xdata"], 20), d["knots"])
    assert np.array_equal(lp.equidistant_knots(d["xdata"], n_param=20), d["knots"])
    assert np.array_equal(o.basis_matrix(d["x"], d["knots"]), d["dense"])
    assert np.array_equal(o.bspline_basis(d["x_out"], d["knots"]), d["masked"])
    assert np.array_equal(o.bspline_basis_deriv(d["x"][:60], d["knots"]), d["deriv"])
    assert np.array_equal(o.bspline_basis_deriv2(d["x"][:60], d["knots"]), d["deriv2"])
    assert bool(d["raised"])  # Basis.bspline rejects x outside the interior knots
    with pytest.raises(ValueError):
        o.basis_matrix(np.array([d["knots"][3] - 0.1]), d["knots"])
    # host mirror's banded evaluation = the 4 non-zeros of the reference's dense rows
    j, vals = lp.banded_basis(d["x"], d["knots"])
    dense = np.zeros((d["x"].shape[0], 21))
    for m in range(4):
        dense[np.arange(d["x"].shape[0]), j - 3 + m] = vals[:, m]
    assert np.array_equal(dense[:, :20], d["dense"])


def test_oracle_trafo_matches_reference_code_bit_for_bit():
    """PTMKnots, BSplineApprox tables, PTMCoef map, PTMSpline over all five regions:
    ptm.py:47-59, 92-265, 298-478; approx.py:360-452."""
    t = golden_io.load("ref_trafo_nparam20.npz")
    kn = o.PTMKnots(-4.0, 4.0, 20)
    assert np.array_equal(kn.knots, t["knots"]) and kn.step == t["knot_step"]
    assert np.array_equal(lp.PTMKnots(-4.0, 4.0, 20).knots, t["knots"])
    sp = o.PTMSpline(kn.knots)
    ap = sp.bspline
    assert np.array_equal(ap.grid, t["grid"]) and ap.step == t["grid_step"]
    assert np.array_equal(ap.basis_grid, t["basis_grid"])
    assert np.array_equal(ap.basis_deriv_grid, t["basis_deriv_grid"])
    assert np.array_equal(ap.basis_deriv2_grid, t["basis_deriv2_grid"])
    assert sp.min_eps == t["min_eps"] and sp.max_eps == t["max_eps"]
    assert np.array_equal(sp.compute_coef(t["delta"]), t["theta"])
    assert np.array_equal(ap.cell(t["r"])[0], t["cell"])
    codes = set()
    for c in range(3):
        z, d = sp.dot_and_deriv_n_fullbatch(t["r"], t["delta"][c])
        assert np.array_equal(z, t["z"][c]) and np.array_equal(d, t["d"][c])
        codes |= set(np.unique(sp.region_code(t["r"])))
    assert codes == {0, 1, 2, 3, 4}
    # the declared tangents of the custom_jvp (approx.py:498-520)
    J = t["theta"].shape[1]
    one, zero = np.ones_like(t["rc"]), np.zeros_like(t["rc"])
    for c in range(3):
        _, (t0, t1) = ap.dot_and_deriv_n_jvp(t["rc"], t["theta"][c], one, np.zeros(J))
        assert np.max(np.abs(t0 - t["dz_dr"][c])) < 1e-14 and np.max(np.abs(t1 - t["dd_dr"][c])) < 1e-13
        for j in range(J):
            e = np.zeros(J)
            e[j] = 1.0
            _, (t0, t1) = ap.dot_and_deriv_n_jvp(t["rc"], t["theta"][c], zero, e)
            assert np.max(np.abs(t0 - t["dz_dtheta"][c][:, j])) < 1e-15
            assert np.max(np.abs(t1 - t["dd_dtheta"][c][:, j])) < 1e-14


def test_setup_code_matches_reference_code():
    """ps() reparameterisations and the trafo Z: penalty.py, constraint.py, gam/var.py:771-860,
    var.py:45-74, 195-224 (same LAPACK in this image, so the null-space bases agree too)."""
    d = golden_io.load("ref_logp_n257.npz")
    for t in range(2):
        ot = o.ps(d[f"x{t}"], 12)
        assert np.array_equal(ot.knots, d[f"knots{t}"])
        assert np.allclose(ot.Z, d[f"Z{t}"], rtol=0, atol=1e-12)
        assert np.allclose(ot.K, d[f"K{t}"], rtol=0, atol=1e-12) and ot.rank == int(d[f"rank{t}"])
        hb = lp.ps(d[f"x{t}"], nbases=12, xname=f"x{t}")
        # the host mirror sums the banded design's columns instead of the dense one; the
        # constraint's null-space basis (eigenvectors of a repeated eigenvalue,
        # constraint.py:58-76) then differs by a rotation R: same column space, K -> R'KR
        R = np.linalg.lstsq(d[f"Z{t}"], hb.Z, rcond=None)[0]
        assert np.allclose(d[f"Z{t}"] @ R, hb.Z, rtol=0, atol=1e-9)
        assert np.allclose(R.T @ d[f"K{t}"] @ R, hb.penalty, rtol=0, atol=1e-9)
        assert abs(abs(np.linalg.det(R)) - 1.0) < 1e-8
        design = d["basis_loc" if t == 0 else "basis_scale"] @ hb.Z
        assert np.allclose(design.sum(axis=0), 0, atol=1e-9)  # test_gam/test_var.py:359-369
    assert np.allclose(o.trafo_Z(20), d["Z_delta"], rtol=0, atol=1e-13)
    assert np.allclose(lp.trafo_reparam(20), d["Z_delta"], rtol=0, atol=1e-13)
    assert np.array_equal(lp.PTMKnots(-4.0, 4.0, 20).knots, d["trafo_knots"])


def test_oracle_logp_grad_information_match_reference_code():
    """LocScaleTransformationDist(batched=False).log_prob per observation (dist.py:185-188,
    339-395, 718-740), MultivariateNormalSingular / Normal priors (gam/dist.py:56-70), the
    forward-mode gradient with the declared rules, and the IWLS information
    (iwls_proposals.py:55-89, 118-144)."""
    d = golden_io.load("ref_logp_n257.npz")
    om = golden_io.oracle_from_fixture(d)
    st = _state(d)
    lp_, g = om.logp_and_grad(st)
    assert rel_err(lp_, d["logp"]) < 1e-14
    for c in range(2):
        ll = om.loglik_terms(st, c)["ll"]
        assert np.max(np.abs(ll - d["loglik_i"][c])) < 1e-13
        assert abs(om.prior(st, c) - d["prior"][c]) < 1e-13 * abs(d["prior"][c])
    go = np.concatenate([g["beta"][0], g["beta"][1], g["delta_z"]], axis=1)
    assert rel_err(go, d["grad"]) < 1e-14
    assert np.max(np.abs(go - d["grad"]) / np.maximum(np.abs(d["grad"]), 1e-3)) < 1e-11
    # the primal's finite differences see the lerp's true slope, not the declared one:
    # agreement only to O(grid step^2) -- recorded, and bounded
    assert 1e-6 < rel_err(d["grad_fd"], d["grad"]) < 1e-3
    # NaN in y => NaN for that observation only (dist.py:340-343)
    om_nan = golden_io.oracle_from_fixture({**{k: d[k] for k in d.files}, "y": d["ynan"]})
    for c in range(2):
        ll = om_nan.loglik_terms(st, c)["ll"]
        assert np.array_equal(np.isnan(ll), np.isnan(d["loglik_i_nan"][c])) and np.isnan(ll).sum() == 1
    # intercept pseudo-parameters: LocationIntercept / LogScaleIntercept.compute_intercept
    # (predictor.py:80-100, 126-130)
    for c in range(2):
        b0, g0 = om.compute_intercepts(st, c)
        assert abs(b0 - d["intercepts"][c, 0]) < 1e-14 and abs(g0 - d["intercepts"][c, 1]) < 1e-14
        b0, g0 = om.compute_intercepts(st, c, sequential=True)
        assert abs(b0 - d["intercepts"][c, 0]) < 1e-14 and abs(g0 - d["intercepts_seq"][c]) < 1e-14
    # observed information: jacfwd(grad(logp)) of each coefficient block through the
    # reference's code on nested forward-mode values (logprob.py:104-113; iwls_proposals.py:
    # 168-184, 227-245, 279-298; kernel.py:240-258), first hess_n observations + priors
    rows = slice(0, int(d["hess_n"]))
    assert np.max(np.abs(d["hess_loc"] - d["hess_loc"].T)) < 1e-12 * np.max(np.abs(d["hess_loc"]))
    for blk, key in ((0, "hess_loc"), (1, "hess_scale"), ("delta_z", "hess_dz")):
        assert rel_err(om.hessian_block(st, 0, blk, rows=rows), d[key]) < 1e-13, key
    _, P0, L0 = om.loc_precision(st, 0)
    _, P1 = om.scale_precision(st, 1)
    assert rel_err(P0, d["precision"][:, 0]) < 1e-14 and rel_err(L0, d["chol"][:, 0]) < 1e-13
    assert rel_err(P1, d["precision"][:, 1]) < 1e-14
    assert rel_err(np.linalg.cholesky(P1), d["chol"][:, 1]) < 1e-13


@pytest.mark.skipif(not os.path.isdir("/root/reference/liesel_ptm"),
                    reason="the reference tree only exists in the build container")
def test_reference_fixtures_are_reproducible(tmp_path):
    """Re-runs the reference's source files over the stand-ins and compares with the
    committed fixtures: they are what that code produces, not hand-edited numbers."""
    import subprocess
    import sys

    script = os.path.join(golden_io.GOLDEN, "make_reference_golden.py")
    env = dict(os.environ, PTM_GOLDEN_OUT=str(tmp_path))
    subprocess.run([sys.executable, script], check=True, env=env, capture_output=True, timeout=600)
    for name in ("ref_basis_nb20.npz", "ref_trafo_nparam20.npz", "ref_logp_n257.npz", "ref_variants_n257.npz"):
        new, old = np.load(tmp_path / name), golden_io.load(name)
        assert sorted(new.files) == sorted(old.files)
        for k in old.files:
            assert np.array_equal(new[k], old[k], equal_nan=True), (name, k)


# ---------------------------------------------------------------------------------------
# the CUDA path vs the reference's code (GPU, through the C ABI)
# ---------------------------------------------------------------------------------------
def _gpu_model(d, dtype=np.float64, y=None):
    dd = {k: d[k] for k in d.files}
    if y is not None:
        dd["y"] = y
    return golden_io.model_from_fixture(dd, dtype=dtype).build()


def _pack(model, d, dtype=np.float64):
    p = d["Z0"].shape[1]
    P = d["params"]
    return model.pack([P[:, :p], P[:, p:2 * p]], P[:, 2 * p:], tau=d["tau"], tau_delta=d["tau_delta"]).astype(dtype)


def _grad_cols(model, d):
    L, _ = model.param_layout()
    p = d["Z0"].shape[1]
    return np.r_[L["beta"][0]:L["beta"][0] + p, L["beta"][1]:L["beta"][1] + p,
                 L["delta_z"]:L["delta_z"] + d["params"].shape[1] - 2 * p]


@pytest.mark.gpu
def test_cuda_basis_vs_reference_code_bit_exact():
    import torch  # noqa: F401

    d = golden_io.load("ref_basis_nb20.npz")
    x = d["x"]
    model = lp.LocScalePTM.new_ptm(np.zeros_like(x), nparam=20, to_float32=False, device="cuda:0")
    Z = np.eye(20)
    model.loc += lp.term.f_ig(lp.Basis(x=x, xname="x", knots=d["knots"], Z=Z, penalty=Z, nbases=20))
    model.build()
    interval, values = model.basis(0)
    interval, values = interval.cpu().numpy(), values.cpu().numpy()
    dense = np.zeros((x.shape[0], 21))
    for m in range(4):
        dense[np.arange(x.shape[0]), interval - 3 + m] = values[:, m]
    assert np.array_equal(dense[:, :20], d["dense"])
    nz = d["dense"] != 0
    first = np.argmax(nz, axis=1)
    assert np.all(interval - 3 <= first) and np.all(first <= interval)


@pytest.mark.gpu
def test_cuda_transform_vs_reference_code():
    import torch

    from tests.helpers import build_case

    t = golden_io.load("ref_trafo_nparam20.npz")
    case = build_case(n=8, n_loc=1, n_scale=0, nbases=8, nparam=20, C=1)
    z, dz = case.model.transform(torch.from_numpy(t["delta"]).cuda(), torch.from_numpy(t["r"]).cuda())
    cell = case.model.last_cell.cpu().numpy()
    sp = o.PTMSpline(t["knots"])
    core = sp.region_code(t["r"]) == 2
    assert np.array_equal(cell[core], t["cell"][core])
    assert np.max(np.abs(z.cpu().numpy() - t["z"])) < 2e-13
    assert np.max(np.abs(dz.cpu().numpy() - t["d"])) < 2e-12


@pytest.mark.gpu
def test_cuda_logp_grad_information_vs_reference_code_f64():
    import torch

    d = golden_io.load("ref_logp_n257.npz")
    model = _gpu_model(d)
    params = torch.from_numpy(_pack(model, d)).cuda()
    logp, grad = model.log_prob_and_grad(params)
    lp_only = model.log_prob(params)
    torch.cuda.synchronize()
    cols = _grad_cols(model, d)
    assert rel_err(logp.cpu().numpy(), d["logp"]) < 1e-12
    assert rel_err(lp_only.cpu().numpy(), d["logp"]) < 1e-12
    assert rel_err(grad.cpu().numpy()[:, cols], d["grad"]) < 1e-12
    P0, L0 = model.precision(0, params)
    P1, L1 = model.precision(1, params)
    assert rel_err(P0.cpu().numpy(), d["precision"][:, 0]) < 1e-11
    assert rel_err(L0.cpu().numpy(), d["chol"][:, 0]) < 1e-9
    assert rel_err(P1.cpu().numpy(), d["precision"][:, 1]) < 1e-11
    assert rel_err(L1.cpu().numpy(), d["chol"][:, 1]) < 1e-9
    # intercept pseudo-parameters from the fused sufficient statistics (ptm_intercept_stats)
    b0, g0 = model.compute_intercepts(params)
    assert np.max(np.abs(b0.cpu().numpy() - d["intercepts"][:, 0])) < 1e-12
    assert np.max(np.abs(g0.cpu().numpy() - d["intercepts"][:, 1])) < 1e-12
    # ... and in the sampler's order, LogScaleIntercept fed the full loc predictor with the
    # updated location intercept (predictor.py:604-609), from the same single sweep
    b0, g0 = model.compute_intercepts(params, sequential=True)
    assert np.max(np.abs(b0.cpu().numpy() - d["intercepts"][:, 0])) < 1e-12
    assert np.max(np.abs(g0.cpu().numpy() - d["intercepts_seq"])) < 1e-12
    # per-observation log-densities through the pointwise sweep
    ll = d["loglik_i"]
    mx = ll.max(axis=0)
    lppd_ref = mx + np.log(np.exp(ll - mx).sum(axis=0)) - np.log(ll.shape[0])
    lppd, pwaic = model.pointwise(params)
    assert rel_err(lppd.cpu().numpy(), lppd_ref) < 1e-12
    assert rel_err(pwaic.cpu().numpy(), ll.var(axis=0)) < 1e-9
    # NaN in y => NaN logp (dist.py:340-343)
    m2 = _gpu_model(d, y=d["ynan"])
    assert np.isnan(m2.log_prob(params).cpu().numpy()).all()


@pytest.mark.gpu
def test_cuda_logp_grad_vs_reference_code_f32():
    """float32 problem: rel 1e-5 (north star) against float64 values on the SAME inputs, i.e.
    the f32-rounded numbers the kernel receives.  The float64 side is the oracle, which the CPU
    part of this file pins bit-for-bit / to 1e-13 on the reference code's fixture; against the
    fixture's own float64 values (inputs NOT rounded) the gradient differs by up to ~3e-5 --
    that is the rounding of y, x, knots, Z to float32, not the kernel -- bounded below too."""
    import torch

    from tests.helpers import pack_grad_layout

    d = golden_io.load("ref_logp_n257.npz")
    model = _gpu_model(d, dtype=np.float32)
    p32 = _pack(model, d, np.float32)
    params = torch.from_numpy(p32).cuda()
    logp, grad = model.log_prob_and_grad(params)
    torch.cuda.synchronize()
    cols = _grad_cols(model, d)
    d32 = {k: (d[k].astype(np.float32).astype(np.float64) if d[k].dtype == np.float64 and d[k].ndim > 0
               else d[k]) for k in d.files}
    om = golden_io.oracle_from_fixture(d32)
    L, P = model.param_layout()
    st = golden_io.state_from_params(om, p32.astype(np.float64), L)
    lp_ref, gd = om.logp_and_grad(st)
    g_ref = pack_grad_layout(L, P, gd)
    g = grad.cpu().numpy()
    assert rel_err(logp.cpu().numpy(), lp_ref) < 1e-5
    for c in range(g.shape[0]):
        assert rel_err(g[c], g_ref[c]) < 1e-5
    assert rel_err(logp.cpu().numpy(), d["logp"]) < 1e-5
    assert rel_err(g[:, cols], d["grad"]) < 1e-4  # unrounded inputs: input-rounding bound
