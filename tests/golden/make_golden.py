#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ from the NumPy oracle.

The reference cannot be imported in this image, so these
vectors come from the oracle (oracle/ptm_oracle.py), which restates the reference
file by file; they pin the oracle against regressions and give the CUDA path fixed
inputs/outputs that travel to the GPU box.  Run from the repo root:

    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ptm_oracle as o  # noqa: E402
from tests.helpers import build_case, oracle_eval  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def golden_basis():
    """(1) interval indices + 4 basis values for x on and around every knot."""
    nb = 20
    knots = o.equidistant_knots(np.array([-2.0, 2.0]), nb)
    xs = []
    for j in range(3, nb + 1):
        k = knots[j]
        for v in (np.nextafter(k, -np.inf), k, np.nextafter(k, np.inf)):
            if knots[3] <= v <= knots[nb]:
                xs.append(v)
    xs += list(np.linspace(knots[3], knots[nb], 57))
    x = np.array(xs)
    D = o.basis_matrix(x, knots)
    interval = (np.searchsorted(knots, x, side="right") - 1).astype(np.int32)
    np.savez(os.path.join(OUT, "basis_nb20.npz"), x=x, knots=knots, dense=D, interval=interval)


def golden_trafo():
    """(2) grid cells / kappa on grid points and region boundaries; (3) theta(delta)."""
    kn = o.PTMKnots(-4.0, 4.0, 20)
    sp = o.PTMSpline(kn.knots)
    g = sp.bspline.grid
    r = np.concatenate([
        g, np.nextafter(g, -np.inf), np.nextafter(g, np.inf),
        [sp.min_eps, sp.max_eps, np.nextafter(sp.min_eps, -np.inf), np.nextafter(sp.max_eps, np.inf)],
        np.linspace(-9, 9, 501),
    ])
    rng = np.random.default_rng(3)
    delta = np.stack([np.zeros(20), rng.normal(size=20), 0.3 * rng.normal(size=20)])
    theta = sp.compute_coef(delta)
    z = np.zeros((3, r.shape[0]))
    d = np.zeros_like(z)
    for c in range(3):
        z[c], d[c] = sp.dot_and_deriv_n_fullbatch(r, delta[c])
    code = sp.region_code(r)
    cell = np.where(code == 2, np.searchsorted(g, r, side="right") - 1, -1).astype(np.int32)
    kappa = np.where(code == 2, (r - g[np.clip(cell, 0, 1001)]) / sp.bspline.step, np.nan)
    np.savez(os.path.join(OUT, "trafo_nparam20.npz"), knots=kn.knots, grid=g, r=r, delta=delta,
             theta=theta, z=z, d=d, code=code.astype(np.int32), cell=cell, kappa=kappa)


def golden_logp():
    """(4) logp + full gradient, incl. NaN in y and far tails; (6) prior-only pieces."""
    cases = {
        "n1_c1": dict(n=1, C=1), "n7_c3": dict(n=7, C=3), "n100_c1": dict(n=100, C=1),
        "n4097_c3": dict(n=4097, C=3),
    }
    for name, kw in cases.items():
        case = build_case(n_loc=2, n_scale=1, nbases=12, nparam=20, build=False, **kw)
        lp, g = oracle_eval(case)
        save_case(name, case, lp, g)

    def with_nan(y):
        y = y.copy()
        y[3] = np.nan
        return y

    case = build_case(n=64, n_loc=1, n_scale=1, nbases=10, C=2, build=False, y_override=with_nan)
    lp, g = oracle_eval(case)
    save_case("nan_in_y", case, lp, g)

    def far(y):
        y = y.copy()
        y[::5] *= 6.0  # deep into both linear tails
        return y

    case = build_case(n=300, n_loc=2, n_scale=1, nbases=12, C=3, build=False, y_override=far)
    lp, g = oracle_eval(case)
    save_case("far_tails", case, lp, g)


def save_case(name, case, lp, g, **extra):
    om = case.omodel
    d = dict(y=om.y, params=case.params, logp=lp, grad=g, nparam=om.nparam, trafo_Z=om.Zd,
             n_terms=len(om.terms))
    for t, term in enumerate(om.terms):
        d[f"x{t}"] = term.x
        d[f"knots{t}"] = term.knots
        d[f"Z{t}"] = term.Z
        d[f"K{t}"] = term.K
        d[f"rank{t}"] = term.rank
        d[f"pred{t}"] = 0 if term.predictor == "loc" else 1
    d.update(extra)
    np.savez(os.path.join(OUT, f"logp_{name}.npz"), **d)


def golden_xtwx():
    """(5) XtWX / precision / Cholesky for p = 19 and p = 29."""
    for nb in (20, 30):
        case = build_case(n=1500, n_loc=1, n_scale=1, nbases=nb, C=2, build=False)
        x, P, L = case.omodel.loc_precision(case.state, 0)
        lp, g = oracle_eval(case)
        save_case(f"xtwx_nb{nb}", case, lp, g, xtwx=x, precision=P, chol=L)


if __name__ == "__main__":
    golden_basis()
    golden_trafo()
    golden_logp()
    golden_xtwx()
    print("golden fixtures written to", OUT)
