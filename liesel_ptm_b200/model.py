"""Host-side mirror of the reference's operator interface for the hot path.

``LocScalePTM`` keeps the reference's construction surface
(model/model.py:279-383: ``LocScalePTM.new_ptm(response, nparam, a, b, ...)``,
``model.loc += term.f_ig(ps(x, nbases=20, xname="x"))``, ``model.build()``) and
exposes what the MCMC kernels call on it:

* ``log_prob`` / ``log_prob_and_grad``     -> logprob.py:36-69, 104-140 (``LogProb``,
  ``FlatLogProb``: log_prob, grad) and iwls_fixed.py:91-116
* ``GaussianLocCholInfo`` / ``GaussianScaleCholInfo`` -> iwls_proposals.py:13-144
* ``basis`` / ``transform``                -> Basis.bspline (bspline/approx.py:116-210),
  dot_and_deriv_n_fullbatch (bspline/util.py:91-103)

All numerics run in the CUDA library through the C ABI (``include/ptm_b200.h``).
torch is used only for device memory and streams.  Chain states travel as one
``[C, P]`` array per call -- the layout ``ravel_pytree`` gives the JAX side.
"""

from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .spline import Basis, PTMKnots, approx_grid, term, trafo_reparam

_NP = {torch.float64: np.float64, torch.float32: np.float32}
_VARIANTS = {"ptm": _lib.PTM_TRAFO_PTM, "onion": _lib.PTM_TRAFO_ONION, "identity": _lib.PTM_TRAFO_IDENTITY}


def _on_device(fn):
    """Run an operator with the model's GPU as the current device (the C ABI launches on the
    current device and refuses a handle that lives elsewhere)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)

    return wrapped


class _Predictor:
    """``model.loc`` / ``model.scale``: supports ``+=`` with a term (predictor.py:23-60)."""

    def __init__(self, kind: int, name: str) -> None:
        self.kind = kind
        self.name = name
        self.terms: dict[str, term] = {}

    def __iadd__(self, other):
        items = other if isinstance(other, (list, tuple)) else [other]
        for t in items:
            if not isinstance(t, term):
                raise TypeError(f"expected a term, got {type(t)}")
            if t.name in self.terms:
                raise RuntimeError(f"{self} already contains a term of name {t.name}.")
            self.terms[t.name] = t
        return self


class LocScalePTM:
    """Location-scale penalized transformation model, B200 hot path."""

    def __init__(self, response, knots, trafo_lambda: float = 0.1, to_float32: bool = True,
                 device: str | torch.device = "cuda", bspline: str = "ptm") -> None:
        """``bspline``: "ptm" (PTMSpline over ``PTMKnots``), "onion" (OnionSpline over
        ``OnionKnots``; experimental in the reference) or "identity" (the Gaussian location-scale
        model), as ``PTMDist`` switches them (model/model.py:98-131, 292)."""
        if bspline not in _VARIANTS:
            raise ValueError(f"bspline must be one of {sorted(_VARIANTS)}, got {bspline!r}")
        self.bspline = bspline
        self.to_float32 = to_float32
        self.dtype = torch.float32 if to_float32 else torch.float64
        self.np_dtype = _NP[self.dtype]
        self.response = np.ascontiguousarray(np.asarray(response), dtype=self.np_dtype)
        self.knots = np.ascontiguousarray(np.asarray(knots), dtype=self.np_dtype)
        # bases of h minus one (ptm.py:55-57) / minus seven (onion.py:50-59)
        self.nparam = self.knots.shape[0] - (11 if bspline == "onion" else 5)
        self.trafo_lambda = float(trafo_lambda)
        # Z_delta is reparameterisation DATA (LAPACK-dependent eigenvector order); a caller
        # that exports it from the host model assigns it before build()
        self.trafo_Z = None
        # PTMCoef(noncentered=...) (var.py:55, 100-107); new_ptm builds it with False
        # (model/model.py:377); set before build() for the non-centred form
        self.trafo_noncentered = False
        self._loc = _Predictor(_lib.PTM_LOC, "$\\mu$")
        self._scale = _Predictor(_lib.PTM_SCALE, "$\\sigma$")
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        self._handle = None
        self._view = None    # observation-shard view handle (ptm_problem_shard_view), if any
        self._owner = True   # this object destroys the problem handle
        self._parent = None  # shard views keep the owning model alive
        self.shard = None
        self._ws = None
        self._ws_chains = 0
        self.is_built = False

    @classmethod
    def new_ptm(cls, response, nparam: int = 20, a: float = -4.0, b: float = 4.0,
                trafo_lambda: float = 0.1, to_float32: bool = True, device="cuda"):
        """Shortcut for convenient model setup (model/model.py:341-383)."""
        dtype = np.float32 if to_float32 else np.float64
        knots = PTMKnots(a, b, nparam=nparam, dtype=dtype)
        return cls(response, knots.knots, trafo_lambda=trafo_lambda, to_float32=to_float32,
                   device=device)

    @classmethod
    def new_gaussian(cls, response, to_float32: bool = True, device="cuda"):
        """Shortcut for a Gaussian location-scale model (model/model.py:386-417): the same knots
        argument as the reference passes, ``bspline="identity"``.  The parameter rows keep the
        (ignored) shape-parameter entries so that the layout is the one of every other model."""
        dtype = np.float32 if to_float32 else np.float64
        return cls(response, np.linspace(-3.0, 3.0, 10).astype(dtype), to_float32=to_float32,
                   device=device, bspline="identity")

    # -- reference surface: model.loc += ..., model.scale += ... ------------------------
    @property
    def loc(self):
        return self._loc

    @loc.setter
    def loc(self, value):
        if value is not self._loc:
            raise ValueError("Cannot overwrite .loc attribute.")

    @property
    def scale(self):
        return self._scale

    @scale.setter
    def scale(self, value):
        if value is not self._scale:
            raise ValueError("Cannot overwrite .scale attribute.")

    # -- build: copy the static data into HBM once ---------------------------------------
    def build(self) -> "LocScalePTM":
        if self.is_built:
            raise ValueError("Graph was already built.")
        lib = _lib.load()
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("liesel_ptm_b200 needs a CUDA device (sm_100a); there is no CPU path")
        torch.cuda.set_device(self.device)
        nd = self.np_dtype
        self.terms = self.all_terms()
        T = len(self.terms)
        keep = []  # keep numpy arrays alive during create

        def arr(a, dt=nd):
            a = np.ascontiguousarray(np.asarray(a), dtype=dt)
            keep.append(a)
            return a

        def ptr_array(arrays):
            pa = (C.c_void_p * max(len(arrays), 1))()
            for i, a in enumerate(arrays):
                pa[i] = a.ctypes.data
            keep.append(pa)
            return C.cast(pa, C.POINTER(C.c_void_p))

        def i32_array(vals):
            a = arr(np.asarray(vals, dtype=np.int32).reshape(-1), np.int32)
            return a.ctypes.data_as(C.POINTER(C.c_int32))

        n = self.response.shape[0]
        for t, _ in self.terms:
            if t.basis.x.shape[0] != n:
                raise ValueError(f"term {t.name}: covariate has {t.basis.x.shape[0]} rows, response {n}")
        d = _lib.PtmProblemDesc()
        d.dtype = _lib.PTM_F64 if self.dtype == torch.float64 else _lib.PTM_F32
        d.n = n
        d.y = arr(self.response).ctypes.data
        d.n_terms = T
        d.x = ptr_array([arr(t.basis.x) for t, _ in self.terms])
        d.nbases = i32_array([t.basis.nbases for t, _ in self.terms] or [0])
        d.ncoef = i32_array([t.basis.ncoef for t, _ in self.terms] or [0])
        d.predictor = i32_array([k for _, k in self.terms] or [0])
        d.knots = ptr_array([arr(t.basis.knots) for t, _ in self.terms])
        d.Z = ptr_array([arr(t.basis.Z) for t, _ in self.terms])
        d.K = ptr_array([arr(t.penalty) for t, _ in self.terms])
        d.rank = i32_array([t.penalty_rank for t, _ in self.terms] or [0])
        d.nparam = self.nparam
        d.trafo_knots = arr(self.knots).ctypes.data
        if self.trafo_Z is None:
            self.trafo_Z = trafo_reparam(self.nparam, nd)
        self.trafo_Z = np.ascontiguousarray(np.asarray(self.trafo_Z, dtype=nd))
        if self.trafo_Z.shape != (self.nparam, self.nparam - 1):
            raise ValueError(f"trafo_Z must have shape ({self.nparam}, {self.nparam - 1})")
        d.trafo_Z = arr(self.trafo_Z).ctypes.data
        d.trafo_lambda = self.trafo_lambda
        self.grid = approx_grid(self.knots)
        d.trafo_grid = arr(self.grid).ctypes.data
        d.noncentered = i32_array([int(t.is_noncentered) for t, _ in self.terms] or [0])
        d.trafo_noncentered = int(self.trafo_noncentered)
        d.trafo_variant = _VARIANTS[self.bspline]
        h = C.c_void_p()
        _lib.check(lib.ptm_problem_create(C.byref(d), C.byref(h)))
        self._handle = h
        self._lib = lib
        self.n = n
        self.num_params = int(lib.ptm_num_params(h))
        T_ = max(T, 1)
        b0, g0, dz, td = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        boff, toff = (C.c_int64 * T_)(), (C.c_int64 * T_)()
        _lib.check(lib.ptm_param_layout(h, C.byref(b0), C.byref(g0), boff, C.byref(dz), toff, C.byref(td)))
        lib_layout = {
            "beta0": int(b0.value), "gamma0": int(g0.value),
            "beta": [int(boff[i]) for i in range(T)], "delta_z": int(dz.value),
            "tau": [int(toff[i]) for i in range(T)], "tau_delta": int(td.value),
        }
        self.layout, P = self.param_layout()
        if lib_layout != self.layout or P != self.num_params:
            raise RuntimeError(f"parameter layout mismatch: {lib_layout} vs {self.layout}")
        self._sfx = "f64" if self.dtype == torch.float64 else "f32"
        self.is_built = True
        return self

    def __del__(self):
        try:
            if self._view is not None:
                self._lib.ptm_problem_destroy(self._view)
                self._view = None
            if self._handle is not None and self._owner:
                self._lib.ptm_problem_destroy(self._handle)
            self._handle = None
        except Exception:
            pass

    # -- position <-> flat parameter rows ------------------------------------------------
    def all_terms(self) -> list[tuple[term, int]]:
        out = [(t, _lib.PTM_LOC) for t in self._loc.terms.values()]
        return out + [(t, _lib.PTM_SCALE) for t in self._scale.terms.values()]

    def param_layout(self) -> tuple[dict, int]:
        """Offsets of one chain's parameter row (see include/ptm_b200.h)."""
        terms = self.all_terms()
        off, beta = 2, []
        for t, _ in terms:
            beta.append(off)
            off += t.basis.ncoef
        dz = off
        tau0 = dz + self.nparam - 1
        layout = {"beta0": 0, "gamma0": 1, "beta": beta, "delta_z": dz,
                  "tau": [tau0 + i for i in range(len(terms))], "tau_delta": tau0 + len(terms)}
        return layout, tau0 + len(terms) + 1

    def pack(self, beta: Sequence[np.ndarray], delta_z, tau=None, tau_delta=None, beta0=None,
             gamma0=None) -> np.ndarray:
        """Assemble a [C, P] parameter array from per-variable arrays (ravel_pytree order)."""
        L, P = self.param_layout()
        nterms = len(L["beta"])
        C_ = np.asarray(delta_z).shape[0]
        out = np.zeros((C_, P), dtype=self.np_dtype)
        out[:, L["beta0"]] = 0.0 if beta0 is None else beta0
        out[:, L["gamma0"]] = 0.0 if gamma0 is None else gamma0
        for t, b in enumerate(beta):
            out[:, L["beta"][t]: L["beta"][t] + b.shape[1]] = b
        out[:, L["delta_z"]: L["delta_z"] + self.nparam - 1] = delta_z
        if nterms:
            out[:, L["tau"][0]: L["tau"][0] + nterms] = 1.0 if tau is None else tau
        out[:, L["tau_delta"]] = 1.0 if tau_delta is None else tau_delta
        return out

    # -- device plumbing -------------------------------------------------------------------
    def _workspace(self, n_chains: int) -> torch.Tensor:
        if self._ws is None or self._ws_chains != n_chains:
            nbytes = int(self._lib.ptm_workspace_bytes(self._h, n_chains))
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_chains = n_chains
        return self._ws

    def _check_params(self, params: torch.Tensor) -> torch.Tensor:
        if not self.is_built:
            raise RuntimeError("call build() first")
        if params.dim() == 1:
            params = params[None, :]
        if params.dim() != 2 or params.shape[1] != self.num_params:
            raise ValueError(f"params must have shape (C, {self.num_params}), got {tuple(params.shape)}")
        if params.dtype != self.dtype or params.device.type != "cuda":
            raise TypeError(f"params must be a CUDA tensor of dtype {self.dtype}")
        return params.contiguous()

    @staticmethod
    def _stream() -> int:
        return torch.cuda.current_stream().cuda_stream

    @property
    def _h(self):
        """The handle the operators run on: the shard view when one is set, else the problem."""
        return self._view if self._view is not None else self._handle

    def _make_view(self, obs_begin: int, obs_end: int, add_priors: bool):
        if not self.is_built:
            raise RuntimeError("call build() first")
        h = C.c_void_p()
        _lib.check(self._lib.ptm_problem_shard_view(self._handle, obs_begin, obs_end, int(add_priors),
                                                    C.byref(h)))
        return h

    def shard_view(self, obs_begin: int, obs_end: int, add_priors: bool = True) -> "LocScalePTM":
        """Observation-axis sharding: a NEW model object over the same device arrays that
        evaluates observations [obs_begin, obs_end) only (``ptm_problem_shard_view``).  The
        model it was made from is untouched; precision / Cholesky entry points raise on a view
        (they are not additive over shards -- see ``parallel.ObsShardedLogProb.precision``)."""
        import copy

        h = self._make_view(obs_begin, obs_end, add_priors)
        v = copy.copy(self)
        v._view, v._owner, v._parent = h, False, self  # the view keeps the owning model alive
        v._ws, v._ws_chains = None, 0
        v.shard = (int(obs_begin), int(obs_end), bool(add_priors))
        for name in ("_pin_in", "_pin_lp", "_pin_g", "_dev_in", "_dev_lp", "_dev_g"):
            v.__dict__.pop(name, None)
        return v

    def set_shard(self, obs_begin: int, obs_end: int, add_priors: bool = True) -> None:
        """In-place convenience over ``ptm_problem_shard_view`` (single-process checks): this
        object's operators run on a view of its problem until the full range is set again.
        The C handles themselves are immutable."""
        h = self._make_view(obs_begin, obs_end, add_priors)
        if self._view is not None:
            self._lib.ptm_problem_destroy(self._view)
        full = obs_begin == 0 and obs_end == self.n and add_priors
        if full:
            self._lib.ptm_problem_destroy(h)
            h = None
        self._view = h
        self.shard = None if full else (int(obs_begin), int(obs_end), bool(add_priors))
        self._ws, self._ws_chains = None, 0

    # -- the operators -----------------------------------------------------------------------
    @_on_device
    def log_prob(self, params: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Model log-density for a batch of chain states [C, P] -> [C]."""
        params = self._check_params(params)
        Cn = params.shape[0]
        ws = self._workspace(Cn)
        logp = out if out is not None else torch.empty(Cn, dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_logp_{self._sfx}")
        _lib.check(f(self._h, params.data_ptr(), Cn, logp.data_ptr(), ws.data_ptr(), ws.numel(),
                     self._stream()))
        return logp

    @_on_device
    def log_prob_and_grad(self, params: torch.Tensor, out_logp: torch.Tensor | None = None,
                          out_grad: torch.Tensor | None = None):
        """value_and_grad of the model log-density: [C, P] -> ([C], [C, P])."""
        params = self._check_params(params)
        Cn = params.shape[0]
        ws = self._workspace(Cn)
        logp = out_logp if out_logp is not None else torch.empty(Cn, dtype=self.dtype, device=self.device)
        grad = out_grad if out_grad is not None else torch.empty_like(params)
        f = getattr(self._lib, f"ptm_logp_grad_{self._sfx}")
        _lib.check(f(self._h, params.data_ptr(), Cn, logp.data_ptr(), grad.data_ptr(), ws.data_ptr(),
                     ws.numel(), self._stream()))
        return logp, grad

    def log_prob_and_grad_host(self, params_host: np.ndarray):
        """Same call with HOST buffers: H2D of the chain states, kernels, D2H of logp+grad."""
        if not hasattr(self, "_pin_in") or self._pin_in.shape != params_host.shape:
            self._pin_in = torch.empty(params_host.shape, dtype=self.dtype).pin_memory()
            self._pin_lp = torch.empty(params_host.shape[0], dtype=self.dtype).pin_memory()
            self._pin_g = torch.empty(params_host.shape, dtype=self.dtype).pin_memory()
            self._dev_in = torch.empty(params_host.shape, dtype=self.dtype, device=self.device)
            self._dev_lp = torch.empty(params_host.shape[0], dtype=self.dtype, device=self.device)
            self._dev_g = torch.empty(params_host.shape, dtype=self.dtype, device=self.device)
        self._pin_in.numpy()[...] = params_host
        self._dev_in.copy_(self._pin_in, non_blocking=True)
        self.log_prob_and_grad(self._dev_in, self._dev_lp, self._dev_g)
        self._pin_lp.copy_(self._dev_lp, non_blocking=True)
        self._pin_g.copy_(self._dev_g, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._pin_lp.numpy(), self._pin_g.numpy()

    @_on_device
    def log_prob_and_grad_block(self, params: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """logp and gradient as ONE contiguous [C, 1 + P] block (column 0 = logp), written in
        place by the epilogue kernel: the payload of the observation-axis all-reduce."""
        params = self._check_params(params)
        Cn = params.shape[0]
        ws = self._workspace(Cn)
        block = out if out is not None else torch.empty((Cn, 1 + self.num_params), dtype=self.dtype,
                                                        device=self.device)
        f = getattr(self._lib, f"ptm_logp_grad_block_{self._sfx}")
        _lib.check(f(self._h, params.data_ptr(), Cn, block.data_ptr(), ws.data_ptr(), ws.numel(),
                     self._stream()))
        return block

    def last_launch_count(self) -> int:
        return int(self._lib.ptm_last_launch_count())

    def profile(self, on: bool) -> None:
        """Measurement hook: record CUDA events around the three kernels of each call."""
        _lib.check(self._lib.ptm_profile_enable(int(on)))

    def last_main_route(self) -> str:
        """Route of the last logp / logp+grad sweep: "cuda-core" (k_main) or "tcgen05"."""
        return ["cuda-core", "tcgen05"][int(self._lib.ptm_last_main_route())]

    def last_xtwx_route(self) -> str:
        """Route of the last information sweep: "banded" (CUDA cores), "tcgen05-w" (tensor-core
        contraction, weights on CUDA cores) or "tcgen05" (scale predictor on the tensor cores too)."""
        return {0: "banded", 1: "tcgen05-w", 2: "tcgen05"}[self._lib.ptm_last_xtwx_route()]

    @property
    def shared_designs(self) -> int:
        """Location / scale term pairs on the same covariate and knots that the sweeps stage
        and gather once (0 = general layout)."""
        return int(self._lib.ptm_problem_shared_designs(self._h)) if self.is_built else 0

    def last_kernel_ms(self) -> tuple[float, float, float]:
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        _lib.check(self._lib.ptm_last_kernel_ms(C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    @_on_device
    def basis(self, term_index: int):
        """Banded design of a term: (interval [n] int32, values [n, 4])."""
        interval = torch.empty(self.n, dtype=torch.int32, device=self.device)
        values = torch.empty((self.n, 4), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_basis_{self._sfx}")
        _lib.check(f(self._h, term_index, interval.data_ptr(), values.data_ptr(), self._stream()))
        return interval, values

    @_on_device
    def transform(self, delta: torch.Tensor, r: torch.Tensor):
        """h(r), h'(r) for raw shape coefficients delta [C, nparam] or [nparam]
        (TransformationSpline.dot_and_deriv_n_fullbatch, bspline/util.py:91-103)."""
        squeeze = delta.dim() == 1
        d2 = (delta[None, :] if squeeze else delta).contiguous()
        if d2.dim() != 2 or d2.shape[1] != self.nparam:
            raise ValueError(f"coef must have shape ({self.nparam},) or (C, {self.nparam})")
        if r.dim() > 1:
            raise TypeError("dot_and_deriv_n_fullbatch expects x of shape (n,) or ()")
        was_scalar = r.dim() == 0
        rv = r.reshape(-1).contiguous()
        Cn, m = d2.shape[0], rv.shape[0]
        z = torch.empty((Cn, m), dtype=self.dtype, device=self.device)
        dz = torch.empty_like(z)
        cell = torch.empty(m, dtype=torch.int32, device=self.device)
        f = getattr(self._lib, f"ptm_transform_{self._sfx}")
        _lib.check(f(self._h, d2.data_ptr(), Cn, rv.data_ptr(), m, z.data_ptr(), dz.data_ptr(),
                     cell.data_ptr(), self._stream()))
        self.last_cell = cell
        if squeeze:
            z, dz = z[0], dz[0]
        if was_scalar:
            z, dz = z[..., 0], dz[..., 0]
        return z, dz

    @_on_device
    def inverse(self, delta: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        """r with h(r) = z for raw shape coefficients delta [C, nparam] or [nparam]
        (TransformationSpline.dot_inverse_n, bspline/util.py:105-122): exact inverse of the
        reference's forward function."""
        squeeze = delta.dim() == 1
        d2 = (delta[None, :] if squeeze else delta).contiguous()
        if d2.dim() != 2 or d2.shape[1] != self.nparam:
            raise ValueError(f"coef must have shape ({self.nparam},) or (C, {self.nparam})")
        if z.dim() > 1:
            raise TypeError("dot_inverse_n expects y of shape (n,) or ()")
        was_scalar = z.dim() == 0
        zv = z.reshape(-1).contiguous()
        Cn, m = d2.shape[0], zv.shape[0]
        r = torch.empty((Cn, m), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_inverse_{self._sfx}")
        _lib.check(f(self._h, d2.data_ptr(), Cn, zv.data_ptr(), m, r.data_ptr(), self._stream()))
        if squeeze:
            r = r[0]
        if was_scalar:
            r = r[..., 0]
        return r

    def _newdata(self, newdata, grid):
        """Device arrays (y_or_grid [m], X [T, m]) for a predictive sweep.  ``newdata`` maps
        covariate names (``basis.xname``) to arrays like the reference's ``newdata`` dicts;
        None means the training covariates.  Values outside a term's interior knots raise
        like ``Basis.bspline`` (approx.py:197-206)."""
        if not self.is_built:
            raise ValueError("Model must be built with .build() first.")
        nd = self.np_dtype
        cols = []
        for t, _ in self.terms:
            if newdata is None:
                x = t.basis.x
            else:
                if t.basis.xname not in newdata:
                    raise KeyError(f"newdata has no entry for covariate {t.basis.xname!r}")
                x = np.asarray(newdata[t.basis.xname], dtype=nd).reshape(-1)
            kn = t.basis.knots
            lo, hi = kn[3], kn[t.basis.nbases]
            # NaN fails the comparison like jnp.min / jnp.max propagate it (approx.py:197-206)
            if x.size and not (x.min() >= lo and x.max() <= hi):
                raise ValueError(f"Values of x are not inside the range of interior knots, [{lo}, {hi}]")
            cols.append(x)
        g = self.response if grid is None else np.asarray(grid, dtype=nd).reshape(-1)
        m = cols[0].shape[0] if cols else g.shape[0]
        if g.shape[0] == 1 and m > 1:
            g = np.full(m, g[0], dtype=nd)
        if any(c.shape[0] != m for c in cols) or g.shape[0] != m:
            raise ValueError("newdata columns and grid must have the same length")
        X = np.stack(cols) if cols else np.zeros((0, m), dtype=nd)
        return (torch.from_numpy(np.ascontiguousarray(g, dtype=nd)).to(self.device),
                torch.from_numpy(np.ascontiguousarray(X, dtype=nd)).to(self.device), m)

    @_on_device
    def summarise_dist(self, samples: torch.Tensor, grid=None, newdata=None,
                       which=("z", "prob", "log_prob", "cdf")) -> dict:
        """z, prob, log_prob, cdf over posterior samples [S, P] at ``grid`` (default: the
        observed response) and ``newdata`` (default: the training covariates) -- each
        [S, m]; LocScalePTM.init_dist + summarise_dist (model/model.py:698-790)."""
        samples = self._check_params(samples)
        g, X, m = self._newdata(newdata, grid)
        S = samples.shape[0]
        ws = torch.empty(max(int(self._lib.ptm_predict_workspace_bytes(self._h, S)), 1),
                         dtype=torch.uint8, device=self.device)
        want_lp = "log_prob" in which or "prob" in which
        out = {}
        z = torch.empty((S, m), dtype=self.dtype, device=self.device) if "z" in which else None
        lpdf = torch.empty((S, m), dtype=self.dtype, device=self.device) if want_lp else None
        cdf = torch.empty((S, m), dtype=self.dtype, device=self.device) if "cdf" in which else None
        f = getattr(self._lib, f"ptm_predict_{self._sfx}")
        ptr = lambda t: 0 if t is None else t.data_ptr()  # noqa: E731
        _lib.check(f(self._h, samples.data_ptr(), S, g.data_ptr(), X.data_ptr(), m, ptr(z), ptr(lpdf),
                     ptr(cdf), ws.data_ptr(), ws.numel(), self._stream()))
        if z is not None:
            out["z"] = z
        if "prob" in which:
            out["prob"] = torch.exp(lpdf)
        if "log_prob" in which:
            out["log_prob"] = lpdf
        if cdf is not None:
            out["cdf"] = cdf
        return out

    @_on_device
    def predict_quantile(self, samples: torch.Tensor, u, newdata=None) -> torch.Tensor:
        """Predictive quantiles [S, m] at probabilities ``u`` (scalar or one per row):
        init_dist(samples, newdata=...).quantile(u) (dist.py:203-209)."""
        samples = self._check_params(samples)
        uv = np.asarray(u, dtype=self.np_dtype).reshape(-1)
        g, X, m = self._newdata(newdata, uv)
        S = samples.shape[0]
        ws = torch.empty(max(int(self._lib.ptm_predict_workspace_bytes(self._h, S)), 1),
                         dtype=torch.uint8, device=self.device)
        yq = torch.empty((S, m), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_predict_quantile_{self._sfx}")
        _lib.check(f(self._h, samples.data_ptr(), S, g.data_ptr(), X.data_ptr(), m, yq.data_ptr(),
                     ws.data_ptr(), ws.numel(), self._stream()))
        return yq

    @_on_device
    def sample(self, samples: torch.Tensor, newdata=None, generator: torch.Generator | None = None,
               u: torch.Tensor | None = None) -> torch.Tensor:
        """One predictive draw per (posterior sample, row): init_dist(samples, newdata=...).sample()
        (TransformationDist._sample_n, dist.py:193-201: u ~ U(eps, 1), then the quantile
        function).  ``u`` [S, m] may be supplied (e.g. jax.random.uniform with the caller's key)
        for a reproducible draw; otherwise it is drawn here with ``generator``.  Returns [S, m]."""
        samples = self._check_params(samples)
        S = samples.shape[0]
        m_rows = self.n if newdata is None else len(np.asarray(next(iter(newdata.values()))))
        eps = float(np.finfo(self.np_dtype).eps)
        if u is None:
            u = eps + (1.0 - eps) * torch.rand((S, m_rows), dtype=self.dtype, device=self.device, generator=generator)
        u = u.to(device=self.device, dtype=self.dtype).contiguous()
        if u.shape != (S, m_rows):
            raise ValueError(f"u must have shape ({S}, {m_rows})")
        _, X, m = self._newdata(newdata, np.full(m_rows, 0.5, dtype=self.np_dtype))
        ws = torch.empty(max(int(self._lib.ptm_predict_workspace_bytes(self._h, S)), 1),
                         dtype=torch.uint8, device=self.device)
        out = torch.empty((S, m), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_predict_sample_{self._sfx}")
        _lib.check(f(self._h, samples.data_ptr(), S, u.data_ptr(), X.data_ptr(), m, out.data_ptr(),
                     ws.data_ptr(), ws.numel(), self._stream()))
        return out

    @_on_device
    def intercept_stats(self, params: torch.Tensor) -> torch.Tensor:
        """[C, 5] sufficient statistics of the intercept pseudo-Gibbs updates: sum w, sum r,
        sum r^2, sum w^2, sum r w (w = exp(-eta_sigma), r = (y - eta_mu) w) over this handle's
        observation range (additive over shards)."""
        params = self._check_params(params)
        Cn = params.shape[0]
        ws = self._workspace(Cn)
        out = torch.empty((Cn, 5), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_intercept_stats_{self._sfx}")
        _lib.check(f(self._h, params.data_ptr(), Cn, out.data_ptr(), ws.data_ptr(), ws.numel(),
                     self._stream()))
        return out

    @_on_device
    def pointwise(self, samples: torch.Tensor):
        """Pointwise predictive records over posterior samples [S, P] (draws x chains
        flattened): (lppd_i [n], p_waic_i [n]) -- EvaluatePTM._lppdi / .waic
        (model/evaluate.py:106-133, 177-236) as one fused sweep."""
        samples = self._check_params(samples)
        S = samples.shape[0]
        nbytes = int(self._lib.ptm_pointwise_workspace_bytes(self._h, S))
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=self.device)
        lppd = torch.zeros(self.n, dtype=self.dtype, device=self.device)
        pwaic = torch.zeros(self.n, dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_pointwise_{self._sfx}")
        _lib.check(f(self._h, samples.data_ptr(), S, lppd.data_ptr(), pwaic.data_ptr(), ws.data_ptr(),
                     ws.numel(), self._stream()))
        return lppd, pwaic

    def waic(self, samples: torch.Tensor) -> dict:
        """The reference's WAIC table (model/evaluate.py:215-236) from ``pointwise``."""
        lppd_i, p_i = self.pointwise(samples)
        lppd_i, p_i = lppd_i.double(), p_i.double()
        elpd_i = lppd_i - p_i
        n = float(self.n)
        lppd, p = float(lppd_i.sum()), float(p_i.sum())
        return {"waic_lppd": lppd, "waic_elpd": lppd - p,
                "waic_se": float(elpd_i.std(unbiased=False) * n ** 0.5), "waic_p": p,
                "waic_deviance": -2.0 * (lppd - p), "n_warning": int((p_i > 4).sum())}

    def compute_intercepts(self, params: torch.Tensor, sequential: bool = False):
        """LocationIntercept.compute_intercept / LogScaleIntercept.compute_intercept
        (predictor.py:80-100, 126-130) for every chain, from one fused sweep: returns
        (beta0_new, gamma0_new).  ``sequential=False``: both evaluated at the current state
        (the reference's two Calc nodes, each seeing the current beta0).  ``sequential=True``:
        gamma0_new as LogScaleIntercept sees it AFTER the location intercept was updated in
        the same sampler sweep (it is wired to the full ``loc`` predictor, predictor.py:604-609)
        -- exact from the same sums, no second sweep."""
        if self.shard is not None and (self.shard[0] != 0 or self.shard[1] != self.n):
            raise RuntimeError("compute_intercepts on an observation-shard view would use partial sums: "
                               "all-reduce intercept_stats() and call parallel.intercepts_from_stats")
        from .parallel import intercepts_from_stats

        return intercepts_from_stats(self.intercept_stats(params), params, self.layout, self.n,
                                     sequential=sequential)

    @_on_device
    def quadform(self, params: torch.Tensor) -> torch.Tensor:
        """beta_t' K_t beta_t for every chain and term, [C, T]: the quadratic form of the
        Gibbs updates of the smoothing variances (gam/kernel.py:29-39, kernel.py:99-111)."""
        params = self._check_params(params)
        Cn = params.shape[0]
        out = torch.empty((Cn, len(self.terms)), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_quadform_{self._sfx}")
        _lib.check(f(self._h, params.data_ptr(), Cn, out.data_ptr(), self._stream()))
        return out

    @_on_device
    def xtwx(self, term_index: int, params: torch.Tensor) -> torch.Tensor:
        params = self._check_params(params)
        Cn = params.shape[0]
        p = self.terms[term_index][0].basis.ncoef
        ws = self._workspace(Cn)
        out = torch.empty((Cn, p, p), dtype=self.dtype, device=self.device)
        f = getattr(self._lib, f"ptm_xtwx_{self._sfx}")
        _lib.check(f(self._h, term_index, params.data_ptr(), Cn, out.data_ptr(), ws.data_ptr(),
                     ws.numel(), self._stream()))
        return out

    @_on_device
    def precision(self, term_index: int, params: torch.Tensor, with_chol: bool = True):
        params = self._check_params(params)
        Cn = params.shape[0]
        p = self.terms[term_index][0].basis.ncoef
        ws = self._workspace(Cn)
        P = torch.empty((Cn, p, p), dtype=self.dtype, device=self.device)
        L = torch.empty_like(P) if with_chol else None
        f = getattr(self._lib, f"ptm_loc_precision_{self._sfx}")
        _lib.check(f(self._h, term_index, params.data_ptr(), Cn, P.data_ptr(),
                     L.data_ptr() if with_chol else None, ws.data_ptr(), ws.numel(), self._stream()))
        return P, L


HESS_RAW, HESS_MAKE_PD, HESS_JITTER_MAKE_PD, HESS_PLUS_IDENTITY, HESS_OR_IDENTITY = range(5)


@_on_device
def _hessian_block(self, block, params: torch.Tensor, mode: int = HESS_RAW, with_chol: bool = True):
    """Observed information ``-jacfwd(grad(log_prob))`` of one coefficient block for every
    chain (``FlatLogProb.hessian``, logprob.py:104-113) with the reference's post-processing
    ``mode`` (include/ptm_b200.h) and its lower Cholesky factor.  ``block`` = term index or
    ``"delta_z"``.  Returns (info [C, d, d], chol [C, d, d] | None)."""
    params = self._check_params(params)
    Cn = params.shape[0]
    if block == "delta_z":
        blk, d = -1, self.nparam - 1
    else:
        blk = int(block)
        if not 0 <= blk < len(self.terms):
            raise IndexError("block: term index or 'delta_z'")
        d = self.terms[blk][0].basis.ncoef
    nbytes = self._lib.ptm_hessian_workspace_bytes(self._h, Cn, blk)
    if nbytes == 0:
        raise _lib.PtmError(2, self._lib.ptm_last_error().decode())  # PTM_ERR_SHAPE
    ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
    info = torch.empty((Cn, d, d), dtype=self.dtype, device=self.device)
    chol = torch.empty_like(info) if with_chol else None
    f = getattr(self._lib, f"ptm_hessian_block_{self._sfx}")
    _lib.check(f(self._h, blk, int(mode), params.data_ptr(), Cn, info.data_ptr(),
                 chol.data_ptr() if with_chol else None, ws.data_ptr(), ws.numel(), self._stream()))
    return info, chol


LocScalePTM.hessian_block = _hessian_block


@_on_device
def _precision_all(self, params: torch.Tensor, with_chol: bool = True):
    """Precision (and Cholesky factor) of every LOCATION term from one sweep: lists of
    [C, p_t, p_t] tensors in model order."""
    params = self._check_params(params)
    Cn = params.shape[0]
    ps_ = [t.basis.ncoef for t, k in self.terms if k == _lib.PTM_LOC]
    total = sum(Cn * p * p for p in ps_)
    ws = self._workspace(Cn)
    P = torch.empty(max(total, 1), dtype=self.dtype, device=self.device)
    L = torch.empty_like(P) if with_chol else None
    f = getattr(self._lib, f"ptm_loc_precision_all_{self._sfx}")
    _lib.check(f(self._h, params.data_ptr(), Cn, P.data_ptr(), L.data_ptr() if with_chol else None,
                 ws.data_ptr(), ws.numel(), self._stream()))
    outP, outL, off = [], [], 0
    for p in ps_:
        outP.append(P[off: off + Cn * p * p].view(Cn, p, p))
        if with_chol:
            outL.append(L[off: off + Cn * p * p].view(Cn, p, p))
        off += Cn * p * p
    return outP, outL


LocScalePTM.precision_all = _precision_all


class FlatLogProb:
    """``FlatLogProb`` of the reference (logprob.py:72-140) over the fused operator:
    ``log_prob(flat_position)``, ``grad(flat_position)`` with flat positions [C, P]."""

    def __init__(self, model: LocScalePTM) -> None:
        self.model = model

    def __call__(self, flat_position: torch.Tensor) -> torch.Tensor:
        return self.log_prob(flat_position)

    def log_prob(self, flat_position: torch.Tensor) -> torch.Tensor:
        return self.model.log_prob(flat_position)

    def grad(self, flat_position: torch.Tensor) -> torch.Tensor:
        return self.model.log_prob_and_grad(flat_position)[1]

    def value_and_grad(self, flat_position: torch.Tensor):
        return self.model.log_prob_and_grad(flat_position)


class GaussianLocCholInfo:
    """iwls_proposals.py:13-89: ``chol_info_fn`` provider for a location term."""

    def __init__(self, model: LocScalePTM, term_index: int) -> None:
        self.model = model
        self.term_index = term_index

    @classmethod
    def from_smooth(cls, smooth: term, model: LocScalePTM):
        for i, (t, _) in enumerate(model.terms):
            if t is smooth:
                return cls(model, i)
        raise ValueError(f"{smooth.name} is not a term of the model")

    def precision(self, params: torch.Tensor) -> torch.Tensor:
        return self.model.precision(self.term_index, params, with_chol=False)[0]

    def chol_info(self, params: torch.Tensor) -> torch.Tensor:
        return self.model.precision(self.term_index, params, with_chol=True)[1]


GaussianScaleCholInfo = GaussianLocCholInfo  # W == 2 is selected by the term's predictor


class _ObservedCholInfo:
    """Shared body of the reference's observed-information providers: the information of one
    block is evaluated ONCE at the state given at construction (``__post_init__`` ->
    ``current_fisher_info``) and ``chol_info(state)`` returns the stored factor."""

    mode = HESS_MAKE_PD

    def __init__(self, model: LocScalePTM, block, params: torch.Tensor) -> None:
        self.model, self.block = model, block
        self._fisher_info_unprocessed = model.hessian_block(block, params, HESS_RAW, with_chol=False)[0]
        self.fisher_info, self.chol_fisher_info = model.hessian_block(block, params, self.mode)

    @classmethod
    def from_smooth(cls, smooth: term, model: LocScalePTM, params: torch.Tensor):
        for i, (t, _) in enumerate(model.terms):
            if t is smooth:
                return cls(model, i, params)
        raise ValueError(f"{smooth.name} is not a term of the model")

    @classmethod
    def from_coef(cls, model: LocScalePTM, params: torch.Tensor):
        return cls(model, "delta_z", params)

    def current_fisher_info(self, params: torch.Tensor):
        return (self.model.hessian_block(self.block, params, HESS_RAW, with_chol=False)[0],
                self.model.hessian_block(self.block, params, self.mode, with_chol=False)[0])

    def chol_info(self, params: torch.Tensor | None = None) -> torch.Tensor:
        return self.chol_fisher_info

    @property
    def nan_in_cholesky_of_unprocessed_finfo(self) -> bool:
        # jnp.linalg.cholesky gives NaN where torch reports a failed pivot
        return bool((torch.linalg.cholesky_ex(self._fisher_info_unprocessed).info != 0).any())


class PTMCholInfoFixed(_ObservedCholInfo):
    """iwls_proposals.py:202-256 (make_pd of the observed information of delta_z)."""


class CholInfo(_ObservedCholInfo):
    """iwls_proposals.py:259-312 (``+ 1e-6 I`` then make_pd, smooth-term coefficients)."""

    mode = HESS_JITTER_MAKE_PD


class PTMCholInfo(_ObservedCholInfo):
    """iwls_proposals.py:145-199: fixed make_pd information, rescaled by the CURRENT
    ``1 / clip(tau_delta, sqrt(eps))^2`` at every call."""

    def precision(self, params: torch.Tensor) -> torch.Tensor:
        lay, _ = self.model.param_layout()
        scale = params[:, lay["tau_delta"]]
        eps = float(np.sqrt(np.finfo(self.model.np_dtype).eps))
        return self.fisher_info * (1.0 / scale.clamp_min(eps) ** 2)[:, None, None]

    def chol_info(self, params: torch.Tensor) -> torch.Tensor:
        L, info = torch.linalg.cholesky_ex(self.precision(params))
        return torch.where((info != 0)[:, None, None], torch.full_like(L, float("nan")), L)


class ObservedCholInfoOrIdentity(_ObservedCholInfo):
    """iwls_proposals.py:315-372: the factor of ``info + 1e-6 I`` at the CURRENT state, the
    identity when it has NaNs."""

    mode = HESS_JITTER_MAKE_PD

    def chol_info(self, params: torch.Tensor) -> torch.Tensor:
        return self.model.hessian_block(self.block, params, HESS_OR_IDENTITY)[1]


def iwls_kernel_chol_info(model: LocScalePTM, block, params: torch.Tensor) -> torch.Tensor:
    """``kernel.IWLSKernel._chol_info`` without a ``chol_info_fn`` (kernel.py:240-258):
    ``cholesky(-hessian + I)`` at the current state, every transition."""
    return model.hessian_block(block, params, HESS_PLUS_IDENTITY)[1]
