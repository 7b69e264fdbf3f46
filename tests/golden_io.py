"""Load a golden logp fixture and rebuild the model / oracle objects from it."""

from __future__ import annotations

import os

import numpy as np

import liesel_ptm_b200 as lp
from oracle import ptm_oracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def model_from_fixture(d, dtype=np.float64, device="cuda:0", n_rows=None, bspline="ptm"):
    """Product model built from the fixture's reparameterisation DATA (Z, K, knots);
    ``n_rows`` keeps the first observations only; ``bspline`` as PTMDist takes it
    (model/model.py:98-131)."""
    if n_rows is not None:
        d = dict(d)
        d["y"] = d["y"][:n_rows]
        for t in range(int(d["n_terms"])):
            d[f"x{t}"] = d[f"x{t}"][:n_rows]
    f32 = np.dtype(dtype) == np.float32
    if bspline == "ptm":
        model = lp.LocScalePTM.new_ptm(d["y"].astype(dtype), nparam=int(d["nparam"]), to_float32=f32, device=device)
    else:
        kn = (lp.OnionKnots if bspline == "onion" else lp.PTMKnots)(-4.0, 4.0, int(d["nparam"]), dtype=dtype)
        model = lp.LocScalePTM(d["y"].astype(dtype), kn.knots, to_float32=f32, device=device, bspline=bspline)
    for t in range(int(d["n_terms"])):
        Z = d[f"Z{t}"].astype(dtype)
        basis = lp.Basis(x=d[f"x{t}"].astype(dtype), xname=f"x{t}", knots=d[f"knots{t}"].astype(dtype),
                         Z=np.ascontiguousarray(Z), penalty=np.ascontiguousarray(d[f"K{t}"].astype(dtype)),
                         nbases=Z.shape[0])
        tm = lp.term.f_ig(basis)
        tm.penalty_rank = int(d[f"rank{t}"])
        if int(d[f"pred{t}"]) == 0:
            model.loc += tm
        else:
            model.scale += tm
    if "trafo_Z" in d:
        model.trafo_Z = d["trafo_Z"].astype(dtype)
    return model


def oracle_from_fixture(d, bspline="ptm"):
    terms = [
        o.term_from_reparam(d[f"x{t}"], d[f"knots{t}"], d[f"Z{t}"], d[f"K{t}"], int(d[f"rank{t}"]),
                            "loc" if int(d[f"pred{t}"]) == 0 else "scale")
        for t in range(int(d["n_terms"]))
    ]
    om = o.PTMModel(d["y"], terms, nparam=int(d["nparam"]), bspline=bspline)
    om.Zd = d["trafo_Z"]
    return om


def state_from_params(om, params, layout):
    T = len(om.terms)
    L = layout
    return o.ChainState(
        beta=[params[:, L["beta"][t]: L["beta"][t] + om.terms[t].p] for t in range(T)],
        tau=params[:, L["tau"][0]: L["tau"][0] + T] if T else np.zeros((params.shape[0], 0)),
        delta_z=params[:, L["delta_z"]: L["delta_z"] + om.nparam - 1],
        tau_delta=params[:, L["tau_delta"]], beta0=params[:, L["beta0"]], gamma0=params[:, L["gamma0"]],
    )
