"""Single-direction forward-mode values for the NumPy stand-in (test infrastructure).

`Dual(p, t)` carries a primal array and one tangent of the same shape through the
reference's code; NumPy ufuncs / functions are intercepted with `__array_ufunc__` /
`__array_function__`.  `custom_jvp` functions apply their DECLARED rule (that is the point:
the reference's gradient is defined by approx.py:498-520, not by the derivative of its
lerp).  Tangent conventions follow jax: comparisons and `searchsorted` see the primal,
`where` selects tangents, `maximum(x, c)` passes the tangent where x > c (half at ties).
"""
import numpy as np


def P(x):
    return x.p if isinstance(x, Dual) else x


def T(x):
    if isinstance(x, Dual):
        return x.t
    return np.zeros(np.shape(x), dtype=np.float64)


def clamp(idx, shape):
    from . import _clamp_index
    return _clamp_index(idx, shape)


class Dual:
    __array_priority__ = 1000.0
    _is_dual = True

    def __init__(self, p, t):
        self.p = np.asarray(p, dtype=np.float64)
        self.t = np.broadcast_to(np.asarray(t, dtype=np.float64), self.p.shape).copy()

    # ---- array-ish surface
    shape = property(lambda s: s.p.shape)
    ndim = property(lambda s: s.p.ndim)
    dtype = property(lambda s: s.p.dtype)
    size = property(lambda s: s.p.size)
    T = property(lambda s: Dual(s.p.T, s.t.T))

    def __len__(self):
        return len(self.p)

    def __repr__(self):
        return f"Dual({self.p!r}, {self.t!r})"

    def __getitem__(self, idx):
        idx = clamp(idx, self.p.shape)
        return Dual(self.p[idx], self.t[idx])

    @property
    def at(self):
        return _At(self)

    def squeeze(self, axis=None):
        return Dual(self.p.squeeze(axis), self.t.squeeze(axis))

    def sum(self, axis=None, keepdims=False):
        return Dual(self.p.sum(axis=axis, keepdims=keepdims), self.t.sum(axis=axis, keepdims=keepdims))

    def reshape(self, *s):
        return Dual(self.p.reshape(*s), self.t.reshape(*s))

    def astype(self, _dt):
        return self

    # ---- arithmetic
    def __neg__(self):
        return Dual(-self.p, -self.t)

    def __pos__(self):
        return self

    def __add__(self, o):
        return Dual(self.p + P(o), self.t + T(o))

    __radd__ = __add__

    def __sub__(self, o):
        return Dual(self.p - P(o), self.t - T(o))

    def __rsub__(self, o):
        return Dual(P(o) - self.p, T(o) - self.t)

    def __mul__(self, o):
        return Dual(self.p * P(o), self.t * P(o) + self.p * T(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        q = self.p / P(o)
        return Dual(q, (self.t - q * T(o)) / P(o))

    def __rtruediv__(self, o):
        q = P(o) / self.p
        return Dual(q, (T(o) - q * self.t) / self.p)

    def __pow__(self, e):
        assert not isinstance(e, Dual)
        return Dual(self.p ** e, e * self.p ** (e - 1) * self.t)

    def __matmul__(self, o):
        return Dual(self.p @ P(o), self.t @ P(o) + self.p @ T(o))

    def __rmatmul__(self, o):
        return Dual(P(o) @ self.p, T(o) @ self.p + P(o) @ self.t)

    def __lt__(self, o):
        return self.p < P(o)

    def __le__(self, o):
        return self.p <= P(o)

    def __gt__(self, o):
        return self.p > P(o)

    def __ge__(self, o):
        return self.p >= P(o)

    def __eq__(self, o):
        return self.p == P(o)

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


class _At:
    def __init__(self, d):
        self.d = d

    def __getitem__(self, idx):
        d = self.d

        class _S:
            @staticmethod
            def set(v):
                p, t = d.p.copy(), d.t.copy()
                p[idx] = P(v)
                t[idx] = T(v)
                return Dual(p, t)

        return _S


def _maximum(a, b):
    pa, pb = P(a), P(b)
    w = np.where(pa > pb, 1.0, np.where(pa == pb, 0.5, 0.0))
    return Dual(np.maximum(pa, pb), w * T(a) + (1.0 - w) * T(b))


def _minimum(a, b):
    pa, pb = P(a), P(b)
    w = np.where(pa < pb, 1.0, np.where(pa == pb, 0.5, 0.0))
    return Dual(np.minimum(pa, pb), w * T(a) + (1.0 - w) * T(b))


def _d(x):
    return x if isinstance(x, Dual) else Dual(x, np.zeros(np.shape(x)))


_UFUNCS = {
    np.add: lambda a, b: _d(a) + b,
    np.subtract: lambda a, b: _d(a) - b,
    np.multiply: lambda a, b: _d(a) * b,
    np.true_divide: lambda a, b: _d(a) / b,
    np.negative: lambda a: -a,
    np.matmul: lambda a, b: _d(a) @ b,
    np.exp: lambda a: Dual(np.exp(a.p), np.exp(a.p) * a.t),
    np.log: lambda a: Dual(np.log(a.p), a.t / a.p),
    np.sqrt: lambda a: Dual(np.sqrt(a.p), 0.5 * a.t / np.sqrt(a.p)),
    np.square: lambda a: a * a,
    np.power: lambda a, e: a ** e,
    np.isnan: lambda a: np.isnan(a.p),
    np.isfinite: lambda a: np.isfinite(a.p),
    np.less: lambda a, b: P(a) < P(b),
    np.less_equal: lambda a, b: P(a) <= P(b),
    np.greater: lambda a, b: P(a) > P(b),
    np.greater_equal: lambda a, b: P(a) >= P(b),
    np.equal: lambda a, b: P(a) == P(b),
    np.maximum: _maximum,
    np.minimum: _minimum,
}


def _where(c, a, b):
    return Dual(np.where(c, P(a), P(b)), np.where(c, T(a), T(b)))


def _clip(x, a_min=None, a_max=None, **kw):
    a_min = kw.get("min", a_min)
    a_max = kw.get("max", a_max)
    if a_min is not None:
        x = _maximum(x, a_min)
    if a_max is not None:
        x = _minimum(x, a_max)
    return x


def _dot(a, b):
    return Dual(np.dot(P(a), P(b)), np.dot(T(a), P(b)) + np.dot(P(a), T(b)))


def _searchsorted(a, v, side="left", sorter=None):
    return np.searchsorted(P(a), P(v), side=side)


def _einsum(sub, *ops, **kw):
    prim = [P(o) for o in ops]
    out_t = 0.0
    for k, o in enumerate(ops):
        if isinstance(o, Dual):
            out_t = out_t + np.einsum(sub, *[o.t if j == k else prim[j] for j in range(len(ops))])
    return Dual(np.einsum(sub, *prim), out_t)


_FUNCS = {
    np.einsum: _einsum,
    np.where: _where,
    np.clip: _clip,
    np.dot: _dot,
    np.searchsorted: _searchsorted,
    np.shape: lambda a: a.p.shape,
    np.ndim: lambda a: a.p.ndim,
    np.size: lambda a: a.p.size,
    np.isnan: lambda a: np.isnan(a.p),
    np.zeros_like: lambda a, **k: np.zeros_like(a.p),
    np.ones_like: lambda a, **k: np.ones_like(a.p),
}

_LINEAR = {np.concatenate, np.stack, np.sum, np.mean, np.diff, np.reshape, np.expand_dims,
           np.squeeze, np.atleast_1d, np.atleast_2d, np.broadcast_to, np.transpose, np.cumsum,
           np.ravel, np.take}


def _map(x, f):
    if isinstance(x, Dual):
        return f(x)
    if isinstance(x, (tuple, list)) and any(isinstance(v, Dual) for v in x):
        return type(x)(f(v) for v in x)
    return x


def _linear(func, args, kwargs):
    """Structure-only / linear functions: apply to primals and to tangents."""
    pa = [_map(a, P) for a in args]
    ta = [_map(a, T) for a in args]
    return Dual(func(*pa, **kwargs), func(*ta, **kwargs))


def logsumexp(a):
    m = np.max(P(a))
    return np.log(np.sum(np.exp(a - m))) + m
