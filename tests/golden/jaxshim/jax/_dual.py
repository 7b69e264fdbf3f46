"""Forward-mode values for the NumPy stand-in (test infrastructure), nestable.

`Dual(p, t, level)` carries a primal array and one tangent of the same shape through the
reference's code; NumPy ufuncs / functions are intercepted with `__array_ufunc__` /
`__array_function__`.  `custom_jvp` functions apply their DECLARED rule (that is the point:
the reference's gradient is defined by approx.py:498-520, not by the derivative of its
lerp).  Tangent conventions follow jax: comparisons and `searchsorted` see the primal,
`where` selects tangents, `maximum(x, c)` passes the tangent where x > c (half at ties).

Nesting (second derivatives, `jacfwd(grad(f))` of logprob.py:104-113): every Dual has a
`level`; the components of a level-L Dual may themselves be Duals of a LOWER level, and a
value of a lower level is a constant for level L (no perturbation confusion).  The innermost
derivative (jax's `grad`, whose JVP trace is created last) gets the HIGHEST level.  When a
`custom_jvp` function meets level-L arguments its rule runs on the level-L primals and
tangents -- which may carry lower-level tangents -- so the rule body, INCLUDING the primal
outputs it recomputes, is differentiated by the outer levels as ordinary code.  That is what
jax does (`JVPTrace.process_custom_jvp_call` uses the rule's own primal outputs).
"""
import numpy as np


def lvl(x):
    return x.level if isinstance(x, Dual) else 0


def _top(*xs):
    return max(lvl(x) for x in xs)


def P(x, level=None):
    """Primal of x at `level` (default: x's own level); lower-level values are constants."""
    L = lvl(x) if level is None else level
    return x.p if (L > 0 and lvl(x) == L) else x


def T(x, level=None):
    L = lvl(x) if level is None else level
    if L > 0 and lvl(x) == L:
        return x.t
    return np.zeros(np.shape(x), dtype=np.float64)


def clamp(idx, shape):
    from . import _clamp_index
    return _clamp_index(idx, shape)


def _arr(x):
    return x if isinstance(x, Dual) else np.asarray(x, dtype=np.float64)


class Dual:
    __array_priority__ = 1000.0
    _is_dual = True

    def __init__(self, p, t, level=1):
        self.level = level
        self.p = _arr(p)
        assert lvl(self.p) < level and lvl(t) < level, "components must be of a lower level"
        t = _arr(t)
        if np.shape(t) != np.shape(self.p):
            if isinstance(t, Dual) or isinstance(self.p, Dual):
                t = t + 0.0 * self.p  # broadcast through the (lower-level) arithmetic
            else:
                t = np.broadcast_to(t, self.p.shape)
        self.t = t if isinstance(t, Dual) else np.array(t, dtype=np.float64, copy=True)

    # ---- array-ish surface
    shape = property(lambda s: np.shape(s.p))
    ndim = property(lambda s: np.ndim(s.p))
    dtype = property(lambda s: np.dtype(np.float64))
    size = property(lambda s: np.size(s.p))
    T = property(lambda s: Dual(s.p.T, s.t.T, s.level))

    def __len__(self):
        return len(self.p)

    def __repr__(self):
        return f"Dual[{self.level}]({self.p!r}, {self.t!r})"

    def __float__(self):
        return float(self.p)

    def __getitem__(self, idx):
        idx = clamp(idx, np.shape(self.p))
        return Dual(self.p[idx], self.t[idx], self.level)

    @property
    def at(self):
        return _At(self)

    def squeeze(self, axis=None):
        return Dual(self.p.squeeze(axis), self.t.squeeze(axis), self.level)

    def sum(self, axis=None, keepdims=False):
        return Dual(self.p.sum(axis=axis, keepdims=keepdims), self.t.sum(axis=axis, keepdims=keepdims),
                    self.level)

    def reshape(self, *s):
        return Dual(self.p.reshape(*s), self.t.reshape(*s), self.level)

    def astype(self, _dt):
        return self

    def copy(self):
        return Dual(self.p.copy(), self.t.copy(), self.level)

    # ---- arithmetic (at the higher of the two levels; the other operand is a constant there)
    def __neg__(self):
        return Dual(-self.p, -self.t, self.level)

    def __pos__(self):
        return self

    def __add__(self, o):
        L = _top(self, o)
        return Dual(P(self, L) + P(o, L), T(self, L) + T(o, L), L)

    __radd__ = __add__

    def __sub__(self, o):
        L = _top(self, o)
        return Dual(P(self, L) - P(o, L), T(self, L) - T(o, L), L)

    def __rsub__(self, o):
        L = _top(self, o)
        return Dual(P(o, L) - P(self, L), T(o, L) - T(self, L), L)

    def __mul__(self, o):
        L = _top(self, o)
        a, b = P(self, L), P(o, L)
        return Dual(a * b, T(self, L) * b + a * T(o, L), L)

    __rmul__ = __mul__

    def __truediv__(self, o):
        L = _top(self, o)
        a, b = P(self, L), P(o, L)
        q = a / b
        return Dual(q, (T(self, L) - q * T(o, L)) / b, L)

    def __rtruediv__(self, o):
        L = _top(self, o)
        a, b = P(o, L), P(self, L)
        q = a / b
        return Dual(q, (T(o, L) - q * T(self, L)) / b, L)

    def __pow__(self, e):
        assert not isinstance(e, Dual)
        return Dual(self.p ** e, e * self.p ** (e - 1) * self.t, self.level)

    def __matmul__(self, o):
        L = _top(self, o)
        a, b = P(self, L), P(o, L)
        return Dual(a @ b, T(self, L) @ b + a @ T(o, L), L)

    def __rmatmul__(self, o):
        L = _top(self, o)
        a, b = P(o, L), P(self, L)
        return Dual(a @ b, T(o, L) @ b + a @ T(self, L), L)

    def __lt__(self, o):
        return _prim(self) < _prim(o)

    def __le__(self, o):
        return _prim(self) <= _prim(o)

    def __gt__(self, o):
        return _prim(self) > _prim(o)

    def __ge__(self, o):
        return _prim(self) >= _prim(o)

    def __eq__(self, o):
        return _prim(self) == _prim(o)

    __hash__ = None

    # ---- NumPy protocols
    def __array_ufunc__(self, ufunc, method, *inputs, **kw):
        if method != "__call__" or kw.get("out") is not None:
            return NotImplemented
        f = _UFUNCS.get(ufunc)
        if f is None:
            raise NotImplementedError(f"Dual: ufunc {ufunc.__name__}")
        return f(*inputs)

    def __array_function__(self, func, types, args, kwargs):
        f = _FUNCS.get(func)
        if f is not None:
            return f(*args, **kwargs)
        if func in _LINEAR:
            return _linear(func, args, kwargs)
        raise NotImplementedError(f"Dual: function {func.__name__}")


def _prim(x):
    """The plain NumPy primal underneath every level (what comparisons / indices see)."""
    while isinstance(x, Dual):
        x = x.p
    return x


class _At:
    def __init__(self, d):
        self.d = d

    def __getitem__(self, idx):
        d = self.d

        class _S:
            @staticmethod
            def set(v):
                L = _top(d, v)
                return Dual(_set(P(d, L), idx, P(v, L)), _set(T(d, L), idx, T(v, L)), L)

        return _S


def _set(base, idx, val):
    """base with base[idx] = val, where either may be a (lower-level) Dual."""
    if isinstance(base, Dual) or isinstance(val, Dual):
        return _lift(base, max(lvl(base), lvl(val))).at[idx].set(val)
    out = np.array(base, dtype=np.float64, copy=True)
    out[idx] = val
    return out


def _lift(x, level):
    """x as a Dual of `level` (zero tangent) when it is of a lower level."""
    if level <= 0 or lvl(x) >= level:
        return x
    return Dual(x, np.zeros(np.shape(x)), level)


def _maximum(a, b):
    L = _top(a, b)
    pa, pb = P(a, L), P(b, L)
    ra, rb = _prim(pa), _prim(pb)
    w = np.where(ra > rb, 1.0, np.where(ra == rb, 0.5, 0.0))
    return Dual(np.maximum(pa, pb), w * T(a, L) + (1.0 - w) * T(b, L), L)


def _minimum(a, b):
    L = _top(a, b)
    pa, pb = P(a, L), P(b, L)
    ra, rb = _prim(pa), _prim(pb)
    w = np.where(ra < rb, 1.0, np.where(ra == rb, 0.5, 0.0))
    return Dual(np.minimum(pa, pb), w * T(a, L) + (1.0 - w) * T(b, L), L)


def _d(x):
    return x if isinstance(x, Dual) else Dual(x, np.zeros(np.shape(x)))


_UFUNCS = {
    np.add: lambda a, b: _d(a) + b if isinstance(a, Dual) else b + a,
    np.subtract: lambda a, b: a - b if isinstance(a, Dual) else b.__rsub__(a),
    np.multiply: lambda a, b: a * b if isinstance(a, Dual) else b * a,
    np.true_divide: lambda a, b: a / b if isinstance(a, Dual) else b.__rtruediv__(a),
    np.negative: lambda a: -a,
    np.matmul: lambda a, b: a @ b if isinstance(a, Dual) else b.__rmatmul__(a),
    np.exp: lambda a: Dual(np.exp(a.p), np.exp(a.p) * a.t, a.level),
    np.log: lambda a: Dual(np.log(a.p), a.t / a.p, a.level),
    np.sqrt: lambda a: Dual(np.sqrt(a.p), 0.5 * a.t / np.sqrt(a.p), a.level),
    np.square: lambda a: a * a,
    np.power: lambda a, e: a ** e,
    np.isnan: lambda a: np.isnan(_prim(a)),
    np.isfinite: lambda a: np.isfinite(_prim(a)),
    np.less: lambda a, b: _prim(a) < _prim(b),
    np.less_equal: lambda a, b: _prim(a) <= _prim(b),
    np.greater: lambda a, b: _prim(a) > _prim(b),
    np.greater_equal: lambda a, b: _prim(a) >= _prim(b),
    np.equal: lambda a, b: _prim(a) == _prim(b),
    np.maximum: _maximum,
    np.minimum: _minimum,
}


def _where(c, a, b):
    L = _top(a, b)
    c = _prim(c)
    return Dual(np.where(c, P(a, L), P(b, L)), np.where(c, T(a, L), T(b, L)), L)


def _clip(x, a_min=None, a_max=None, **kw):
    a_min = kw.get("min", a_min)
    a_max = kw.get("max", a_max)
    if a_min is not None:
        x = _maximum(x, a_min)
    if a_max is not None:
        x = _minimum(x, a_max)
    return x


def _dot(a, b):
    L = _top(a, b)
    pa, pb = P(a, L), P(b, L)
    return Dual(np.dot(pa, pb), np.dot(T(a, L), pb) + np.dot(pa, T(b, L)), L)


def _searchsorted(a, v, side="left", sorter=None):
    return np.searchsorted(_prim(a), _prim(v), side=side)


def _einsum(sub, *ops, **kw):
    L = _top(*ops)
    prim = [P(o, L) for o in ops]
    out_t = 0.0
    for k, o in enumerate(ops):
        if lvl(o) == L:
            out_t = out_t + np.einsum(sub, *[o.t if j == k else prim[j] for j in range(len(ops))])
    return Dual(np.einsum(sub, *prim), out_t, L)


_FUNCS = {
    np.einsum: _einsum,
    np.where: _where,
    np.clip: _clip,
    np.dot: _dot,
    np.searchsorted: _searchsorted,
    np.shape: lambda a: np.shape(a.p),
    np.ndim: lambda a: np.ndim(a.p),
    np.size: lambda a: np.size(a.p),
    np.isnan: lambda a: np.isnan(_prim(a)),
    np.zeros_like: lambda a, **k: np.zeros_like(_prim(a)),
    np.ones_like: lambda a, **k: np.ones_like(_prim(a)),
}

_LINEAR = {np.concatenate, np.stack, np.sum, np.mean, np.diff, np.reshape, np.expand_dims,
           np.squeeze, np.atleast_1d, np.atleast_2d, np.broadcast_to, np.transpose, np.cumsum,
           np.ravel, np.take}


def _levels_in(x):
    if isinstance(x, Dual):
        return x.level
    if isinstance(x, (tuple, list)):
        return max([_levels_in(v) for v in x] or [0])
    return 0


def _map(x, f):
    if isinstance(x, Dual):
        return f(x)
    if isinstance(x, (tuple, list)) and _levels_in(x) > 0:
        return type(x)(f(v) for v in x)
    return x


def _linear(func, args, kwargs):
    """Structure-only / linear functions: apply to primals and to tangents (of the top level;
    lower levels are handled by the recursive call on the components)."""
    L = max(_levels_in(a) for a in args)
    pa = [_map(a, lambda v: P(v, L)) for a in args]
    ta = [_map(a, lambda v: T(v, L)) for a in args]
    return Dual(func(*pa, **kwargs), func(*ta, **kwargs), L)


def logsumexp(a):
    m = np.max(_prim(a))
    return np.log(np.sum(np.exp(a - m))) + m
