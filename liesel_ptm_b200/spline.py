"""Host-side model set-up: knots, P-spline bases, penalties, constraints.

Mirrors the reference's set-up API for the hot path (same names, argument
meaning and error behaviour): ``equidistant_knots`` / ``PTMKnots``
(bspline/approx.py:11-68, bspline/ptm.py:14-59), ``Penalty.pspline``
(penalty.py:9-15), ``mixed_model`` and ``LinearConstraintEVD`` (constraint.py:9-23,
58-115), ``ps`` and ``term.f_ig`` (gam/var.py:925-1022, 288-383).

This runs once per model on the host (as it does in the reference, which also
builds knots, eigendecompositions and constraint matrices outside the sampling
loop).  Unlike the reference it never materialises the dense (n, p) design
matrix: the constraint only needs the column sums of the banded basis, and the
kernels rebuild the four non-zero basis values per observation on the fly.
"""

from __future__ import annotations

import dataclasses
from typing import Literal

import numpy as np


def _linspace(start, stop, num, dtype):
    """jnp.linspace semantics (jax 0.8.2): start*(1-s) + stop*s, endpoint appended."""
    dt = np.dtype(dtype).type
    start, stop = dt(start), dt(stop)
    if num == 1:
        return np.array([start], dtype=dtype)
    s = (np.arange(num - 1, dtype=dtype) / dt(num - 1)).astype(dtype)
    body = start * (dt(1) - s) + stop * s
    return np.concatenate([body, [stop]]).astype(dtype)


def equidistant_knots(x, n_param: int, order: int = 3, eps: float = 0.01, dtype=np.float64):
    """Equidistant knots for a B-spline (bspline/approx.py:11-68)."""
    if order < 0:
        raise ValueError(f"Invalid {order=}.")
    if n_param < order:
        raise ValueError(f"{n_param=} must not be smaller than {order=}.")
    dt = np.dtype(dtype).type
    x = np.asarray(x, dtype=dtype)
    a, b = x.min(), x.max()
    range_ = b - a
    min_k = a - range_ * dt(eps / 2)
    max_k = b + range_ * dt(eps / 2)
    internal = _linspace(min_k, max_k, n_param - order + 1, dtype)
    step = internal[1] - internal[0]
    left = _linspace(min_k - step * dt(order), min_k - step, order, dtype)
    right = _linspace(max_k + step, max_k + step * dt(order), order, dtype)
    return np.concatenate([left, internal, right]).astype(dtype)


class PTMKnots:
    """Knots for the monotone transformation spline (bspline/ptm.py:14-59)."""

    def __init__(self, a: float, b: float, nparam: int, order: int = 3, eps: float = 0.0,
                 dtype=np.float64) -> None:
        self.a, self.b, self.nparam, self.order = a, b, nparam, order
        self.knots = equidistant_knots(
            np.array([a, b], dtype=dtype), order=order, n_param=nparam + 1, eps=eps, dtype=dtype
        )
        self.step = np.diff(self.knots).mean()


LogIncKnots = PTMKnots


class OnionKnots:
    """Knots of the onion spline (bspline/onion.py:13-59): nparam free increments on [a, b], three
    fixed increments of one knot step on either side, nparam + 7 bases."""

    def __init__(self, a: float, b: float, nparam: int, order: int = 3, dtype=np.float64) -> None:
        self.a, self.b, self.nparam, self.order = a, b, nparam, order
        m = nparam + 5
        self.step = (self.b - self.a) / (m - 3)
        k1 = a - self.step
        k2 = b + self.step
        self.knots = equidistant_knots(
            np.array([k1, k2], dtype=dtype), order=order, n_param=nparam + 7, eps=0.0, dtype=dtype
        )


def approx_grid(knots, ngrid: int = 1000):
    """BSplineApprox.grid (bspline/approx.py:372-380): [min-step, linspace, max+step]."""
    knots = np.asarray(knots)
    dt = knots.dtype.type
    lo, hi = knots[3], knots[-4]
    grid = _linspace(lo, hi, ngrid, knots.dtype)
    step = (hi - lo) / dt(ngrid)
    return np.concatenate([[lo - step], grid, [hi + step]]).astype(knots.dtype)


class Penalty:
    @staticmethod
    def pspline(nparam: int, random_walk_order: int = 2, dtype=np.float64):
        """(nparam x nparam) P-spline penalty D'D (penalty.py:9-15)."""
        D = np.diff(np.identity(nparam, dtype=dtype), random_walk_order, axis=0)
        return D.T @ D


def mixed_model(penalty):
    """Penalty diagonalisation (constraint.py:9-23)."""
    penalty = np.asarray(penalty)
    evalues, evectors = np.linalg.eigh(penalty)
    evalues, evectors = evalues[::-1], evectors[:, ::-1]
    rank = int(np.linalg.matrix_rank(penalty))
    if evectors[0, 0] < 0:
        evectors = -evectors
    scale = np.ones_like(evalues)
    scale[:rank] = evalues[:rank]
    return evectors * (1.0 / np.sqrt(scale))[None, :]


class LinearConstraintEVD:
    """constraint.py:57-115."""

    @staticmethod
    def general(constraint):
        A = np.asarray(constraint)
        nconstraints = A.shape[0]
        AtA = A.T @ A
        _, evecs = np.linalg.eigh(AtA)
        signs = np.sign(evecs[0, :])
        signs = np.where(signs == 0, 1.0, signs)
        evecs = evecs * signs
        rank = int(np.linalg.matrix_rank(AtA))
        Abar = evecs[:, :-rank].T
        return np.linalg.inv(np.r_[A, Abar])[:, nconstraints:]

    @classmethod
    def sumzero_coef(cls, ncoef: int, dtype=np.float64):
        return cls.general(np.ones((1, ncoef), dtype=dtype))

    @classmethod
    def sumzero_term_from_colsums(cls, colsums):
        """sumzero_term (constraint.py:110-115) given 1'B instead of the dense basis."""
        return cls.general(np.asarray(colsums)[None, :])


def banded_basis(x, knots):
    """Knot interval j (knots[j] <= x < knots[j+1]) and the four non-zero cubic
    B-spline values of columns j-3..j: the banded form of Basis.bspline
    (bspline/approx.py:74-210).  Raises like the reference when x leaves the
    interior knot range."""
    x = np.atleast_1d(np.asarray(x))
    k = np.sort(np.asarray(knots, dtype=x.dtype))
    nb = k.shape[0] - 4
    lo, hi = k[3], k[nb]
    if not (x.min() >= lo and x.max() <= hi):
        raise ValueError(
            f"Values of x are not inside the range of interior knots, [{lo}, {hi}]"
        )
    j = np.clip(np.searchsorted(k, x, side="right") - 1, 3, nb)
    kk = lambda o: k[j + o]  # noqa: E731
    d10 = kk(1) - kk(0)
    n1a = (kk(1) - x) / d10
    n1b = (x - kk(0)) / d10
    d2a, d2b = kk(1) - kk(-1), kk(2) - kk(0)
    n2a = (kk(1) - x) / d2a * n1a
    n2b = (x - kk(-1)) / d2a * n1a + (kk(2) - x) / d2b * n1b
    n2c = (x - kk(0)) / d2b * n1b
    d3a, d3b, d3c = kk(1) - kk(-2), kk(2) - kk(-1), k[np.minimum(j + 3, nb + 3)] - kk(0)
    vals = np.stack(
        [
            (kk(1) - x) / d3a * n2a,
            (x - kk(-2)) / d3a * n2a + (kk(2) - x) / d3b * n2b,
            (x - kk(-1)) / d3b * n2b + (k[np.minimum(j + 3, nb + 3)] - x) / d3c * n2c,
            (x - kk(0)) / d3c * n2c,
        ],
        axis=1,
    )
    return j.astype(np.int32), vals


@dataclasses.dataclass
class Basis:
    """What ``ps()`` leaves behind for the kernels (gam/var.py Basis, 622-690)."""

    x: np.ndarray
    xname: str
    knots: np.ndarray
    Z: np.ndarray          # (nbases, p): product of all reparameterisations
    penalty: np.ndarray    # (p, p) after the reparameterisations
    nbases: int            # B-spline bases before constraints

    @property
    def ncoef(self) -> int:
        return self.Z.shape[1]


def ps(
    x,
    nbases: int,
    xname: str | None = None,
    knots=None,
    order: int = 3,
    outer_ok: bool = False,
    penalty=None,
    diagonalize_penalty: bool = True,
    scale_penalty: bool = True,
    constraint: Literal["sumzero_term", "sumzero_coef", "custom"] | None = "sumzero_term",
    constraint_matrix=None,
    dtype=np.float64,
) -> Basis:
    """P-spline basis (gam/var.py:925-1022), same defaults as the reference."""
    if not xname:
        raise ValueError("Either x must be a named lsl.Var, or xname must be set.")
    if order != 3:
        raise ValueError("the B200 path implements cubic B-splines (order=3) only")
    if outer_ok:
        raise ValueError("outer_ok=True is not supported by the B200 path")
    x = np.asarray(x, dtype=dtype)
    if knots is None:
        knots = equidistant_knots(x, n_param=nbases, order=order, dtype=dtype)
    knots = np.asarray(knots, dtype=dtype)
    j, vals = banded_basis(x, knots)  # raises ValueError outside the knot range
    K = Penalty.pspline(nbases, 2, dtype) if penalty is None else np.asarray(penalty, dtype=dtype)
    Z = np.eye(nbases, dtype=dtype)
    if scale_penalty:  # Basis.reparam_scale_penalty, gam/var.py:806-828
        K = K / np.linalg.norm(K, ord=np.inf)
    if diagonalize_penalty:  # Basis.reparam_diagonalize_penalty, 771-804
        Z1 = mixed_model(K)
        K = Z1.T @ K @ Z1
        Z = Z @ Z1
    if constraint is not None:  # Basis.reparam_constraint, 862-922
        if constraint == "sumzero_term":
            colsums = np.zeros(nbases + 1, dtype=np.float64)
            for m in range(4):
                colsums += np.bincount(j - 3 + m, weights=vals[:, m], minlength=nbases + 1)
            Zc = LinearConstraintEVD.sumzero_term_from_colsums(colsums[:nbases] @ Z)
        elif constraint == "sumzero_coef":
            Zc = LinearConstraintEVD.sumzero_coef(Z.shape[1], dtype)
        elif constraint == "custom":
            if constraint_matrix is None:
                raise ValueError("If type_='custom', `constraint_matrix` must be passed.")
            Zc = np.asarray(constraint_matrix, dtype=dtype)
        else:
            raise ValueError(f"constraint {constraint!r} is not supported by the B200 path")
        K = Zc.T @ K @ Zc
        Z = Z @ Zc
    return Basis(x=x, xname=xname, knots=knots, Z=np.ascontiguousarray(Z.astype(dtype)),
                 penalty=np.ascontiguousarray(K.astype(dtype)), nbases=nbases)


class term:
    """Smooth term ``basis @ coef`` with a rank-deficient Gaussian prior
    (gam/var.py:121-170).  ``term.f_ig`` mirrors gam/var.py:288-383; the
    inverse-gamma hyper-prior on the variance is an O(1)-per-chain scalar that
    stays in the host sampler, so only its default parameters are recorded."""

    def __init__(self, basis: Basis, name: str, ig_concentration: float = 1.0,
                 ig_scale: float = 0.005) -> None:
        self.basis = basis
        self.name = name
        self.penalty = basis.penalty
        self.penalty_rank = int(np.linalg.matrix_rank(basis.penalty))  # gam/var.py:147
        self.nbases = basis.ncoef
        self.ig_concentration = ig_concentration
        self.ig_scale = ig_scale
        self.is_noncentered = False

    def reparam_noncentered(self) -> "term":
        """gam/var.py:176-197: ``coef ~ N(0, scale^2 inv(penalty))`` becomes ``latent ~ N(0,
        inv(penalty)); coef = scale * latent``.  The chain state then holds the latent
        coefficient; must be called before ``LocScalePTM.build()``."""
        self.is_noncentered = True
        return self

    @classmethod
    def f_ig(cls, basis: Basis, fname: str = "f", ig_concentration: float = 1.0,
             ig_scale: float = 0.005, noncentered: bool = False) -> "term":
        t = cls(basis, name=f"{fname}({basis.xname})", ig_concentration=ig_concentration,
                ig_scale=ig_scale)
        if noncentered:  # gam/var.py:283-284, 379-380
            t.reparam_noncentered()
        return t


def lin(x, xname: str | None = None, add_intercept: bool = False, penalty=None, dtype=np.float64) -> Basis:
    """Linear basis (gam/var.py:1025-1061, Basis.new_linear 721-770): design ``[x]`` or ``[1, x]``.

    The kernels evaluate every term through the banded cubic B-spline design, and cubic
    B-splines reproduce linear functions exactly: with the Greville abscissae
    ``g_j = (k_{j+1} + k_{j+2} + k_{j+3}) / 3`` of a four-basis knot vector over the range of x,
    ``sum_j B_j(x) g_j = x`` and ``sum_j B_j(x) = 1``.  So the linear design is the spline
    design with ``Z = [g]`` (or ``[1, g]``) -- no extra kernel path; parity to rounding (1e-15).
    ``penalty`` is the (p, p) prior penalty of the wrapping ``term`` (identity by default, i.e.
    an iid normal prior, the way the reference's notebooks wrap ``lin``)."""
    if not xname:
        raise ValueError("Either x must be a named lsl.Var, or xname must be set.")
    x = np.asarray(x, dtype=dtype)
    if x.ndim != 1:
        raise ValueError("the B200 path takes one covariate column per linear term")
    nb = 4
    knots = equidistant_knots(x, n_param=nb, order=3, dtype=dtype)
    banded_basis(x, knots)  # range check, raises like the reference
    k = knots.astype(np.float64)
    grev = np.array([(k[j + 1] + k[j + 2] + k[j + 3]) / 3.0 for j in range(nb)])
    Z = np.stack([np.ones(nb), grev], axis=1) if add_intercept else grev[:, None]
    p = Z.shape[1]
    K = np.eye(p) if penalty is None else np.asarray(penalty, dtype=np.float64).reshape(p, p)
    return Basis(x=x, xname=xname, knots=knots, Z=np.ascontiguousarray(Z.astype(dtype)),
                 penalty=np.ascontiguousarray(K.astype(dtype)), nbases=nb)


def trafo_reparam(nparam: int, dtype=np.float64):
    """Z_delta of PTMCoef.new_rw1_sumzero + PTMCoef.__init__ (var.py:164-224, 45-74):
    delta = Z_delta @ delta_z with an iid Normal(0, tau_delta) prior on delta_z."""
    pen = Penalty.pspline(nparam, 1, dtype)
    pen = pen / np.linalg.norm(pen, ord=np.inf)
    Z = mixed_model(pen)[:, :-1]
    pen2 = np.eye(nparam - 1, dtype=dtype)
    pen2 = pen2 / np.linalg.norm(pen2, ord=np.inf)
    return np.ascontiguousarray((Z @ mixed_model(pen2)).astype(dtype))
