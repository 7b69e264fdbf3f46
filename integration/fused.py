"""JAX side of the drop-in boundary: the fused PTM operator behind ``jax.ffi`` + ``custom_vjp``.

This module is what a liesel-ptm maintainer adds next to ``liesel_ptm/model/model.py`` to run
the hot path through ``libptm_b200.so``.  It needs jax, liesel and liesel_ptm -- none of which
exist in the build image of this repo (Python 3.12, no jax wheel) -- so it is imported lazily
and NEVER from ``liesel_ptm_b200`` itself; ``tests/test_host.py`` byte-compiles it and checks
the FFI target table against the symbols ``csrc/ptm_xla_ffi.cc`` defines.  All numerics live
behind the C ABI and are tested there (ctypes, ``tests/``).

Seams (reference file:line under ``liesel_ptm/``):

* model interface -- ``gs.LieselInterface(self.graph)`` is created at ``model/model.py:477``
  and installed with ``eb.set_model(self.interface)`` at ``model/model.py:1369``.  Every goose
  kernel touches the model only through ``extract_position / update_state / log_prob``
  (``logprob.py:50-69,104-140``, ``iwls_fixed.py:91-116,159-186``).  ``FusedPTMInterface``
  keeps that protocol; its ``update_state`` writes the small coefficient vectors into the
  state and evaluates ``_model_log_prob`` with ONE fused sweep instead of re-running the
  graph's n-vector nodes.
* IWLS information -- ``gs.IWLSKernel(kernel_kwargs={"chol_info_fn": fn})`` at
  ``model/model.py:1001,1038,1089,1126,1208``; providers in ``iwls_proposals.py:13-144``.
  ``FusedLocCholInfo`` has the same ``from_smooth(...) / chol_info(model_state)`` surface.
* observed information -- ``CholInfo`` / ``PTMCholInfoFixed`` / ``kernel.IWLSKernel._chol_info``
  (``iwls_proposals.py:168-364``, ``kernel.py:240-258``): ``FusedObservedCholInfo``.

Chains: goose vmaps the kernels over the chain axis.  Every ``ffi_call`` below is registered
with ``vmap_method="expand_dims"`` on a ``[1, P]`` operand, so the batched call arrives in the
handler as ONE ``[C, 1, P]`` -> reshaped ``[C, P]`` block: the chain batch of the kernel.
"""

from __future__ import annotations

import ctypes
import dataclasses
import os
from typing import Any, Callable, Sequence

import numpy as np

# (target name registered with XLA, exported symbol of libptm_xla_ffi.so) per precision suffix
FFI_TARGETS = {
    "ptm_logp": "PtmLogp",
    "ptm_logp_grad": "PtmLogpGrad",
    "ptm_logp_grad_block": "PtmLogpGradBlock",
    "ptm_xtwx": "PtmXtwx",
    "ptm_loc_precision": "PtmLocPrecision",
    "ptm_loc_precision_all": "PtmLocPrecisionAll",
    "ptm_intercept_stats": "PtmInterceptStats",
    "ptm_quadform": "PtmQuadform",
    "ptm_hessian_block": "PtmHessianBlock",
}
SUFFIX = {"float32": "F32", "float64": "F64"}

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "liesel_ptm_b200")


def register_ffi_targets(ffi_lib_path: str | None = None) -> None:
    """``jax.ffi.register_ffi_target`` for every handler of ``csrc/ptm_xla_ffi.cc``."""
    import jax

    lib = ctypes.CDLL(ffi_lib_path or os.path.join(_PKG, "libptm_xla_ffi.so"))
    for name, sym in FFI_TARGETS.items():
        for dt, sfx in SUFFIX.items():
            jax.ffi.register_ffi_target(f"{name}_{dt}", jax.ffi.pycapsule(getattr(lib, sym + sfx)),
                                        platform="CUDA")


# ------------------------------------------------------------------------------------------
# problem set-up: the static data of a built reference model -> ptm_problem_create
# ------------------------------------------------------------------------------------------
def reparam_matrix(basis, knots, order: int = 3) -> np.ndarray:
    """Z with ``basis.value == B_raw(x) @ Z``: the product of the reparameterisations the
    reference applied to the basis function (``gam/var.py:771-860``: diagonalisation, sum-to-zero
    constraint).  They are closed over inside ``basis.value_node.function``, so Z is recovered
    from the function itself on a probe grid that covers every knot interval (exact: the map
    is linear and B_raw has full column rank there)."""
    from liesel_ptm.bspline.approx import basis_matrix

    k = np.asarray(knots, dtype=np.float64)
    lo, hi = k[order], k[-order - 1]
    probe = np.linspace(lo, hi, 8 * (k.shape[0] - 1) + 1)
    raw = np.asarray(basis_matrix(probe, k, order, False), dtype=np.float64)
    final = np.asarray(basis.value_node.function(probe), dtype=np.float64)
    Z, *_ = np.linalg.lstsq(raw, final, rcond=None)
    return Z


@dataclasses.dataclass
class FusedProblem:
    """Owns the ``PtmProblem*`` of one built ``liesel_ptm.LocScalePTM`` and the name -> offset
    layout of the flat ``[P]`` parameter row (``include/ptm_b200.h``)."""

    handle: int
    dtype: Any
    num_params: int
    names: dict  # var name -> (offset, size) in the flat row
    term_index: dict  # smooth name -> term index
    lib: Any

    @classmethod
    def from_model(cls, model, term_knots: dict, lib_path: str | None = None) -> "FusedProblem":
        """``model``: a built reference ``LocScalePTM``.  ``term_knots``: smooth name -> knots of
        its ``ps()`` basis (the reference does not keep them on the ``Basis`` object; use
        ``fused_ps`` below, which records them)."""
        import jax.numpy as jnp

        from liesel_ptm_b200 import _lib as abi  # ctypes declarations of include/ptm_b200.h

        if lib_path:
            os.environ["PTM_B200_LIB"] = lib_path
        lib = abi.load()
        dt = np.float32 if model.to_float32 else np.float64
        keep, terms = [], []
        for kind, pred in ((abi.PTM_LOC, model.loc), (abi.PTM_SCALE, model.scale)):
            for name, t in pred.terms.items():
                if not hasattr(t, "basis") or name not in term_knots:
                    continue  # intercepts; other term types stay in the graph (out of scope)
                terms.append((name, kind, t))

        def arr(a, d=dt):
            a = np.ascontiguousarray(np.asarray(a), dtype=d)
            keep.append(a)
            return a

        def ptrs(arrays):
            pa = (ctypes.c_void_p * max(len(arrays), 1))(*[a.ctypes.data for a in arrays])
            keep.append(pa)
            return ctypes.cast(pa, ctypes.POINTER(ctypes.c_void_p))

        def i32(vals):
            return arr(np.asarray(vals, dtype=np.int32), np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int32))

        d = abi.PtmProblemDesc()
        d.dtype = abi.PTM_F32 if dt == np.float32 else abi.PTM_F64
        y = arr(model.response.value)
        d.n, d.y, d.n_terms = y.shape[0], y.ctypes.data, len(terms)
        Zs = [arr(reparam_matrix(t.basis, term_knots[name])) for name, _, t in terms]
        d.x = ptrs([arr(t.basis.x.value) for _, _, t in terms])
        d.nbases = i32([Z.shape[0] for Z in Zs] or [0])
        d.ncoef = i32([Z.shape[1] for Z in Zs] or [0])
        d.predictor = i32([k for _, k, _ in terms] or [0])
        d.knots = ptrs([arr(term_knots[name]) for name, _, _ in terms])
        d.Z = ptrs(Zs)
        d.K = ptrs([arr(t.coef.dist_node["penalty"].value) for _, _, t in terms])
        d.rank = i32([int(np.linalg.matrix_rank(np.asarray(t.coef.dist_node["penalty"].value)))
                      for _, _, t in terms] or [0])
        coef = model.trafo  # PTMCoef (var.py:45-119): knots, Z, latent_coef
        # PTMDist(bspline=...) (model/model.py:98-131): the dist node's partial class tells the variant
        pdc = getattr(model.response.dist_node, "partial_dist_class", None)
        bs_inst = getattr(pdc, "keywords", {}).get("bspline", None)
        variant = 1 if type(bs_inst).__name__ == "OnionSpline" else (
            2 if getattr(getattr(pdc, "func", None), "__name__", "") == "GaussianPseudoTransformationDist" else 0)
        d.trafo_variant = variant
        d.nparam = int(np.asarray(coef.knots).shape[0] - (11 if variant == 1 else 5))
        d.trafo_knots = arr(coef.knots).ctypes.data
        d.trafo_Z = arr(coef.Z).ctypes.data
        d.trafo_lambda = float(model.trafo_lambda)
        d.trafo_grid = arr(coef.bspline.grid).ctypes.data if hasattr(coef, "bspline") else None
        h = ctypes.c_void_p()
        abi.check(lib.ptm_problem_create(ctypes.byref(d), ctypes.byref(h)))

        # name -> (offset, size): the ravel order of include/ptm_b200.h
        T = len(terms)
        b0, g0, dz, td = (ctypes.c_int64() for _ in range(4))
        boff, toff = (ctypes.c_int64 * max(T, 1))(), (ctypes.c_int64 * max(T, 1))()
        abi.check(lib.ptm_param_layout(h, ctypes.byref(b0), ctypes.byref(g0), boff, ctypes.byref(dz), toff,
                                       ctypes.byref(td)))
        names = {model.loc.loc_intercept.name: (int(b0.value), 1),
                 model.scale.log_scale_intercept.name: (int(g0.value), 1),
                 coef.latent_coef.name: (int(dz.value), d.nparam - 1),
                 coef.latent_coef.dist_node["scale"].name: (int(td.value), 1)}
        for i, (name, _, t) in enumerate(terms):
            names[t.coef.name] = (int(boff[i]), int(Zs[i].shape[1]))
            names[t.scale.name] = (int(toff[i]), 1)
        del jnp
        return cls(handle=int(h.value), dtype=dt, num_params=int(lib.ptm_num_params(h)), names=names,
                   term_index={name: i for i, (name, _, _) in enumerate(terms)}, lib=lib)

    def workspace_bytes(self, n_chains: int) -> int:
        return int(self.lib.ptm_workspace_bytes(ctypes.c_void_p(self.handle), n_chains))

    def close(self) -> None:
        self.lib.ptm_problem_destroy(ctypes.c_void_p(self.handle))


def fused_ps(x, nbases: int, xname: str, knots_out: dict, **kwargs):
    """``gam.var.ps`` that records the knots it used under ``xname`` (for FusedProblem)."""
    from liesel_ptm.bspline.approx import equidistant_knots
    from liesel_ptm.gam.var import ps

    knots = kwargs.pop("knots", None)
    if knots is None:
        knots = equidistant_knots(np.asarray(x), n_param=nbases, order=kwargs.get("order", 3))
    knots_out[xname] = np.asarray(knots)
    return ps(x, nbases=nbases, xname=xname, knots=knots, **kwargs)


# ------------------------------------------------------------------------------------------
# the operators: ffi_call + custom_vjp
# ------------------------------------------------------------------------------------------
def make_operators(prob: FusedProblem, max_chains: int = 4) -> dict[str, Callable]:
    """Differentiable ``logp(flat[P])`` and the information operators, vmappable over chains.

    ``max_chains``: the chain batch the engine will vmap over (``num_chains`` of ``run_mcmc``).
    XLA sizes result buffers at trace time, so the device workspace every call receives is
    ``ptm_workspace_bytes(prob, max_chains)``.  On the float32 tensor-core route that includes
    the dl/deta exchange buffer (8 bytes per observation and chain: 8.2 GB at n = 1M x 1024
    chains), so pass the real chain count rather than a generous bound."""
    import jax
    import jax.numpy as jnp

    dt = jnp.float32 if prob.dtype == np.float32 else jnp.float64
    tag = "float32" if prob.dtype == np.float32 else "float64"
    P = prob.num_params
    # XLA sizes result buffers at trace time: the workspace is sized for the largest chain batch
    ws = jax.ShapeDtypeStruct((prob.workspace_bytes(max_chains),), jnp.uint8)
    handle = np.int64(prob.handle)

    def call(name, outs):
        # attributes (handle, term, block, mode) are passed as keyword arguments of the call
        return jax.ffi.ffi_call(f"{name}_{tag}", (*outs, ws), vmap_method="expand_dims")

    def value_and_grad_rows(rows):  # rows [1, P] (one chain; [C, 1, P] under vmap)
        lp, g, _ = call("ptm_logp_grad", (jax.ShapeDtypeStruct((1,), dt), jax.ShapeDtypeStruct((1, P), dt)))(
            rows, handle=handle)
        return lp[0], g[0]

    @jax.custom_vjp
    def logp(flat):
        lp, _ = call("ptm_logp", (jax.ShapeDtypeStruct((1,), dt),))(flat[None, :], handle=handle)
        return lp[0]

    def logp_fwd(flat):
        lp, g = value_and_grad_rows(flat[None, :])
        return lp, g  # the fused sweep already produced the gradient: stash it

    def logp_bwd(g, ct):
        return (ct * g,)

    logp.defvjp(logp_fwd, logp_bwd)

    def loc_precision(flat, term: int, p: int):
        Pm, L, _ = call("ptm_loc_precision", (jax.ShapeDtypeStruct((1, p, p), dt),) * 2)(
            flat[None, :], handle=handle, term=np.int64(term))
        return Pm[0], L[0]

    def hessian_block(flat, block: int, size: int, mode: int):
        Hm, L, _ = call("ptm_hessian_block", (jax.ShapeDtypeStruct((1, size, size), dt),) * 2)(
            flat[None, :], handle=handle, block=np.int64(block), mode=np.int64(mode))
        return Hm[0], L[0]

    def intercept_stats(flat):
        st, _ = call("ptm_intercept_stats", (jax.ShapeDtypeStruct((1, 5), dt),))(flat[None, :], handle=handle)
        return st[0]

    return {"logp": logp, "value_and_grad": lambda f: value_and_grad_rows(f[None, :]),
            "loc_precision": loc_precision, "hessian_block": hessian_block, "intercept_stats": intercept_stats}


# ------------------------------------------------------------------------------------------
# seam 1: the model interface
# ------------------------------------------------------------------------------------------
def make_interface(graph, prob: FusedProblem, ops: dict):
    """``FusedPTMInterface(gs.LieselInterface)``: drop-in for ``model.interface``
    (``model/model.py:477``); install with ``model.interface = make_interface(...)`` before
    ``run_mcmc`` (``eb.set_model(self.interface)``, ``model/model.py:1369``)."""
    import jax.numpy as jnp
    import liesel.goose as gs

    order = sorted(prob.names.items(), key=lambda kv: kv[1][0])

    class FusedPTMInterface(gs.LieselInterface):
        def _flat(self, model_state):
            return jnp.concatenate([jnp.atleast_1d(model_state[k].value).ravel().astype(prob.dtype)
                                    for k, _ in order])

        def _hyperprior_log_prob(self, model_state):
            # O(1)-per-chain priors that are not part of the fused density: InverseGamma on
            # tau_t^2 (gam/var.py:352-360), Weibull . Exp . SoftClip on tau_delta^2
            # (var.py:638-683): read from their (cheap, scalar) dist nodes in the state
            return sum(jnp.sum(model_state[name].value) for name in self._hyper_nodes)

        _hyper_nodes = tuple(n for n in graph.nodes if n.endswith("_log_prob") and "tau" in n)

        def update_state(self, position, model_state):
            # write the small coefficient vectors; skip the graph's n-vector nodes entirely
            new = dict(model_state)
            for key, value in position.items():
                new[key] = new[key]._replace(value=value, outdated=False)
            lp = ops["logp"](self._flat(new)) + self._hyperprior_log_prob(new)
            new["_model_log_prob"] = new["_model_log_prob"]._replace(value=lp, outdated=False)
            return new

        def log_prob(self, model_state):
            return model_state["_model_log_prob"].value

    return FusedPTMInterface(graph)


# ------------------------------------------------------------------------------------------
# seam 2: chol_info_fn providers (iwls_proposals.py)
# ------------------------------------------------------------------------------------------
@dataclasses.dataclass
class FusedLocCholInfo:
    """``GaussianLocCholInfo`` / ``GaussianScaleCholInfo`` surface (iwls_proposals.py:13-144):
    ``from_smooth(smooth, model)`` then ``chol_info(model_state) -> (p, p)`` lower Cholesky."""

    interface: Any
    ops: dict
    term: int
    p: int

    @classmethod
    def from_smooth(cls, smooth, model, prob: FusedProblem, ops: dict):
        return cls(interface=model, ops=ops, term=prob.term_index[smooth.name],
                   p=prob.names[smooth.coef.name][1])

    def precision(self, model_state):
        return self.ops["loc_precision"](self.interface._flat(model_state), self.term, self.p)[0]

    def chol_info(self, model_state):
        return self.ops["loc_precision"](self.interface._flat(model_state), self.term, self.p)[1]


@dataclasses.dataclass
class FusedObservedCholInfo:
    """Observed information of one coefficient block: ``-jacfwd(grad)`` + the reference's
    post-processing, selected by ``mode`` (include/ptm_b200.h, ptm_hessian_block_*):
    1 = make_pd (PTMCholInfo / PTMCholInfoFixed / IWLSKernelFixed.init_state,
    iwls_proposals.py:168-184,227-245, iwls_fixed.py:118-143), 2 = + 1e-6 I then make_pd
    (CholInfo / ObservedCholInfoOrIdentity, 279-298, 333-364), 3 = cholesky(-H + I)
    (kernel.IWLSKernel._chol_info, kernel.py:240-258)."""

    interface: Any
    ops: dict
    block: int
    size: int
    mode: int = 2

    def fisher_info(self, model_state):
        return self.ops["hessian_block"](self.interface._flat(model_state), self.block, self.size, self.mode)[0]

    def chol_info(self, model_state):
        return self.ops["hessian_block"](self.interface._flat(model_state), self.block, self.size, self.mode)[1]


def install(model, term_knots: dict, num_chains: int = 4) -> FusedProblem:
    """Everything a ``run_mcmc`` call needs: build the problem from a built reference model,
    register the FFI targets, swap the interface.  Usage::

        knots = {}
        model.loc += term.f_ig(fused_ps(x, nbases=20, xname="x", knots_out=knots))
        model.build()
        fused.install(model, knots, num_chains=4)
        results = model.run_mcmc(seed=1, warmup=300, posterior=500, num_chains=4)
    """
    register_ffi_targets()
    prob = FusedProblem.from_model(model, term_knots)
    ops = make_operators(prob, max_chains=num_chains)
    model.interface = make_interface(model.graph, prob, ops)
    model._fused = (prob, ops)
    return prob


def all_target_symbols() -> Sequence[str]:
    """Exported handler symbols this module expects in libptm_xla_ffi.so."""
    return [sym + sfx for sym in FFI_TARGETS.values() for sfx in SUFFIX.values()]
