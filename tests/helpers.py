"""Shared test plumbing: build the same seeded case for the CUDA path and the oracle."""

from __future__ import annotations

import dataclasses

import numpy as np

import liesel_ptm_b200 as lp
from oracle import ptm_oracle as o


@dataclasses.dataclass
class Case:
    model: lp.LocScalePTM
    omodel: o.PTMModel
    state: o.ChainState
    params: np.ndarray
    layout: dict
    P: int


def rel_err(a, b):
    """max |a-b| / max |b| (per the parity definition in SURVEY.md section 7)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if np.isnan(b).any() or np.isnan(a).any():
        if not np.array_equal(np.isnan(a), np.isnan(b)):
            return np.inf
        m = ~np.isnan(b)
        a, b = a[m], b[m]
        if a.size == 0:
            return 0.0
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def build_case(n=500, n_loc=2, n_scale=1, nbases=12, nparam=20, C=3, dtype=np.float64, seed=0,
               build=True, y_override=None, amp=0.1, beta0=0.05, gamma0=-0.03, tau=None,
               nbases_list=None, device="cuda:0", bspline="ptm") -> Case:
    y, X = o.synth_data(max(n, 64), n_loc, n_scale, seed=seed, dtype=dtype)
    y, X = y[:n], X[:, :n]
    if y_override is not None:
        y = y_override(y)
    T = n_loc + n_scale
    nbl = nbases_list or [nbases] * T
    if bspline == "ptm":
        model = lp.LocScalePTM.new_ptm(y, nparam=nparam, to_float32=(np.dtype(dtype) == np.float32),
                                       device=device)
    else:
        kn = (lp.OnionKnots if bspline == "onion" else lp.PTMKnots)(-4.0, 4.0, nparam, dtype=dtype)
        model = lp.LocScalePTM(y, kn.knots, to_float32=(np.dtype(dtype) == np.float32), device=device,
                               bspline=bspline)
    oterms = []
    for t in range(T):
        # knots from the covariate's design range (not the sample) so that n = 1 works
        knots = lp.equidistant_knots(np.array([-2.0, 2.0]), n_param=nbl[t], dtype=dtype)
        basis = lp.ps(X[t], nbases=nbl[t], xname=f"x{t}", knots=knots, dtype=dtype)
        tm = lp.term.f_ig(basis, fname="f" if t < n_loc else "g")
        if t < n_loc:
            model.loc += tm
        else:
            model.scale += tm
        oterms.append(o.term_from_reparam(X[t], basis.knots, basis.Z, basis.penalty, tm.penalty_rank,
                                          "loc" if t < n_loc else "scale"))
    omodel = o.PTMModel(y, oterms, nparam=nparam, dtype=dtype, bspline=bspline)
    omodel.Zd = lp.trafo_reparam(nparam, dtype)  # same reparameterisation data on both sides
    st = o.synth_state(oterms, nparam, C, seed=seed + 1, dtype=dtype, amp=amp)
    st.beta0[:] = beta0
    st.gamma0[:] = gamma0
    rng = np.random.default_rng(seed + 2)
    st.tau[:] = rng.uniform(0.6, 1.5, size=st.tau.shape) if tau is None else tau
    st.tau_delta[:] = rng.uniform(0.4, 0.9, size=C)
    layout, P = model.param_layout()
    params = model.pack(st.beta, st.delta_z, tau=st.tau, tau_delta=st.tau_delta, beta0=st.beta0,
                        gamma0=st.gamma0)
    if build:
        model.build()
    return Case(model=model, omodel=omodel, state=st, params=params, layout=layout, P=P)


def pack_grad(case: Case, g: dict) -> np.ndarray:
    L = case.layout
    C = g["delta_z"].shape[0]
    out = np.zeros((C, case.P), dtype=np.float64)
    out[:, L["beta0"]] = g["beta0"]
    out[:, L["gamma0"]] = g["gamma0"]
    for t, gb in enumerate(g["beta"]):
        out[:, L["beta"][t]: L["beta"][t] + gb.shape[1]] = gb
    out[:, L["delta_z"]: L["delta_z"] + g["delta_z"].shape[1]] = g["delta_z"]
    for t in range(len(g["beta"])):
        out[:, L["tau"][t]] = g["tau"][:, t]
    out[:, L["tau_delta"]] = g["tau_delta"]
    return out


def oracle_eval(case: Case):
    """logp [C] and packed gradient [C, P] from the NumPy oracle."""
    lp_, g = case.omodel.logp_and_grad(case.state)
    return lp_, pack_grad(case, g)


def oracle_from_workload(wl, chain_idx=None, as_f64=False):
    """Oracle model + chain states for a ``bench_workload.Workload`` (the bench's own seeded
    data and states).  ``chain_idx`` selects a subset of the chain batch (chains are
    independent); ``as_f64`` evaluates the f64 oracle on the workload's (f32-rounded) numbers."""
    from tests import golden_io

    dt = np.float64 if as_f64 else wl.params.dtype
    terms = [
        o.term_from_reparam(wl.X[i].astype(dt), t.basis.knots.astype(dt), t.basis.Z.astype(dt),
                            t.penalty.astype(dt), t.penalty_rank, "loc" if k == 0 else "scale")
        for i, (t, k) in enumerate(wl.model.all_terms())
    ]
    om = o.PTMModel(wl.y.astype(dt), terms, nparam=wl.model.nparam, dtype=dt)
    om.spline = o.PTMSpline(wl.model.knots.astype(dt))
    Zd = wl.model.trafo_Z if wl.model.trafo_Z is not None else lp.trafo_reparam(wl.model.nparam, wl.params.dtype)
    om.Zd = np.asarray(Zd).astype(dt)
    L, P = wl.model.param_layout()
    params = wl.params if chain_idx is None else wl.params[np.asarray(chain_idx)]
    st = golden_io.state_from_params(om, params.astype(dt), L)
    return om, st, L, P


def pack_grad_layout(L, P, g: dict) -> np.ndarray:
    C = g["delta_z"].shape[0]
    out = np.zeros((C, P), dtype=np.float64)
    out[:, L["beta0"]] = g["beta0"]
    out[:, L["gamma0"]] = g["gamma0"]
    for t, gb in enumerate(g["beta"]):
        out[:, L["beta"][t]: L["beta"][t] + gb.shape[1]] = gb
        out[:, L["tau"][t]] = g["tau"][:, t]
    out[:, L["delta_z"]: L["delta_z"] + g["delta_z"].shape[1]] = g["delta_z"]
    out[:, L["tau_delta"]] = g["tau_delta"]
    return out
