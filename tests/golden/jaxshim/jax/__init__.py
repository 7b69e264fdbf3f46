"""NumPy stand-in for the subset of jax the reference's hot-path files use.

Test infrastructure (see ../README.md).  Semantics follow jax with
``jax_enable_x64=True``: float64 everywhere, out-of-bounds scalar reads clamp, out-of-bounds
``.at[i].set`` writes are dropped, ``lax.switch`` clamps its index.
"""
from __future__ import annotations

import functools
import types
import typing

import numpy as _np
import scipy.special as _sps


class Array(_np.ndarray):
    """ndarray with jax's ``.at`` property and clamped scalar reads."""

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        return _wrap(_np.ndarray.__getitem__(self, _clamp_index(idx, self.shape)))


def _clamp_one(i, n):
    if isinstance(i, (int, _np.integer)):
        i = int(i)
        if i < 0:
            i += n
        return min(max(i, 0), n - 1)
    if isinstance(i, _np.ndarray) and i.dtype.kind in "iu":
        j = _np.where(i < 0, i + n, i)
        return _np.clip(j, 0, n - 1).view(_np.ndarray)
    return i


def _clamp_index(idx, shape):
    """jax gather semantics: integer (array) indices wrap once, then clamp."""
    if isinstance(idx, tuple):
        if any(i is Ellipsis or i is None for i in idx):
            return idx
        return tuple(_clamp_one(i, shape[ax]) if ax < len(shape) else i
                     for ax, i in enumerate(idx))
    if len(shape) >= 1:
        return _clamp_one(idx, shape[0])
    return idx


class _At:
    def __init__(self, a):
        self.a = a

    def __getitem__(self, idx):
        return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx):
        self.a, self.idx = a, idx

    def _oob(self):
        i = self.idx
        if isinstance(i, (int, _np.integer)) and self.a.ndim >= 1:
            n = self.a.shape[0]
            return not (-n <= int(i) < n)
        return False

    def set(self, v):
        out = _np.array(self.a, copy=True)
        if not self._oob():
            out[self.idx] = v
        return out.view(Array)

    def add(self, v):
        out = _np.array(self.a, copy=True)
        if not self._oob():
            _np.add.at(out, self.idx, v)
        return out.view(Array)


def _isdual(x):
    return getattr(x, "_is_dual", False)


def _wrap(x):
    if isinstance(x, Array):
        return x
    if isinstance(x, _np.ndarray):
        return x.view(Array)
    if isinstance(x, tuple):
        return tuple(_wrap(v) for v in x)
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x


def _wrapping(fn):
    @functools.wraps(fn)
    def inner(*a, **k):
        return _wrap(fn(*a, **k))

    return inner


# --------------------------------------------------------------------------- jax.numpy
def _linspace(start, stop, num=50, endpoint=True):
    """jax's arithmetic: ``start*(1-s) + stop*s`` with ``s = iota(div)/div`` and the
    endpoint appended (jax/_src/numpy/lax_numpy.py, restated from the published source)."""
    start = _np.float64(start)
    stop = _np.float64(stop)
    if num == 1:
        return _np.array([start]).view(Array)
    div = num - 1 if endpoint else num
    s = _np.arange(div, dtype=_np.float64) / _np.float64(div)
    out = start * (1.0 - s) + stop * s
    if endpoint:
        out = _np.concatenate([out, [stop]])
    return out.view(Array)


def _vectorize(pyfunc=None, *, excluded=frozenset(), signature=None):
    def deco(f):
        vf = _np.vectorize(f, excluded=set(excluded), signature=signature)
        return _wrapping(vf)

    return deco if pyfunc is None else deco(pyfunc)


class _IndexTrick:
    def __init__(self, obj):
        self.obj = obj

    def __getitem__(self, item):
        if isinstance(item, tuple) and any(_isdual(x) for x in item):
            # r_[(scalar, dual vector)]: numpy's index trick would iterate the Dual forever
            assert self.obj is _np.r_
            return _np.concatenate([x if _isdual(x) else _np.atleast_1d(x) for x in item])
        return _wrap(self.obj[item])


def _clip(x, min=None, max=None, **kw):
    if "a_min" in kw:
        min = kw["a_min"]
    if "a_max" in kw:
        max = kw["a_max"]
    if _isdual(x):
        from ._dual import _clip as dclip
        return dclip(x, min, max)
    return _wrap(_np.clip(x, min, max))


class _Module(types.ModuleType):
    def __init__(self, name, base, overrides=None):
        super().__init__(name)
        self._base = base
        for k, v in (overrides or {}).items():
            setattr(self, k, v)

    def __getattr__(self, name):
        v = getattr(self._base, name)
        if callable(v) and not isinstance(v, type):
            v = _wrapping(v)
        setattr(self, name, v)
        return v


_linalg = _Module("jax.numpy.linalg", _np.linalg)
numpy = _Module(
    "jax.numpy",
    _np,
    dict(
        linspace=_linspace,
        vectorize=_vectorize,
        clip=_clip,
        r_=_IndexTrick(_np.r_),
        c_=_IndexTrick(_np.c_),
        linalg=_linalg,
        ndarray=Array,
        array=lambda x, dtype=None: x if _isdual(x) else _np.array(x, dtype=dtype).view(Array),
        asarray=lambda x, dtype=None: x if _isdual(x) else _np.asarray(x, dtype=dtype).view(Array),
    ),
)


# --------------------------------------------------------------------------- transforms
def jit(fn=None, **_kw):
    if fn is None:
        return lambda f: f
    return fn


def _tree_index(t, i):
    if isinstance(t, (tuple, list)):
        return type(t)(_tree_index(v, i) for v in t)
    return t[i]


def _tree_stack(items):
    first = items[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_tree_stack([it[j] for it in items]) for j in range(len(first)))
    if first is None:
        return None
    if any(_isdual(v) for v in items):
        from ._dual import Dual, P, T, lvl
        L = max(lvl(v) for v in items)
        return Dual(_tree_stack([P(v, L) for v in items]), _tree_stack([T(v, L) for v in items]), L)
    return _np.stack([_np.asarray(v) for v in items]).view(Array)


def vmap(fn, in_axes=0, out_axes=0):
    assert out_axes == 0

    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                assert ax == 0
                leaf = a
                while isinstance(leaf, (tuple, list)):
                    leaf = leaf[0]
                n = _np.shape(leaf)[0]
                break
        outs = []
        for i in range(n):
            call = [a if ax is None else _tree_index(a if (isinstance(a, (tuple, list)) or _isdual(a)) else _wrap(_np.asarray(a)), i)
                    for a, ax in zip(args, axes)]
            outs.append(fn(*call))
        return _tree_stack(outs)

    return mapped


class custom_jvp:
    """Calls the primal; the declared rule stays reachable as ``.jvp``."""

    def __init__(self, fn, nondiff_argnums=()):
        self.fn = fn
        self.jvp = None
        functools.update_wrapper(self, fn)

    def defjvp(self, rule):
        self.jvp = rule
        return rule

    def __call__(self, *a, **k):
        if any(_isdual(v) for v in a):
            # the rule runs on the primals / tangents of the HIGHEST level present; they may
            # carry lower-level tangents, which then see the rule body as ordinary code
            from ._dual import Dual, P, T, lvl
            L = max(lvl(v) for v in a)
            as_arr = lambda x: x if _isdual(x) else _wrap(_np.asarray(x))  # noqa: E731
            prim = tuple(as_arr(P(v, L)) for v in a)
            tang = tuple(as_arr(T(v, L)) for v in a)
            out_p, out_t = self.jvp(prim, tang)
            if isinstance(out_p, (tuple, list)):
                return type(out_p)(Dual(p_, t_, L) for p_, t_ in zip(out_p, out_t))
            return Dual(out_p, out_t, L)
        return self.fn(*a, **k)


class _Lax(types.ModuleType):
    @staticmethod
    def fori_loop(lo, hi, body, init):
        val = init
        for i in range(int(lo), int(hi)):
            val = body(i, val)
        return val

    @staticmethod
    def cond(pred, true_fn, false_fn, *operands):
        return true_fn(*operands) if bool(pred) else false_fn(*operands)

    @staticmethod
    def switch(index, branches, *operands):
        i = min(max(int(index), 0), len(branches) - 1)
        return branches[i](*operands)

    @staticmethod
    def scan(f, init, xs, length=None):
        n = length
        if n is None:
            leaf = xs
            while isinstance(leaf, (tuple, list)):
                leaf = leaf[0]
            n = _np.shape(leaf)[0]
        carry, ys = init, []
        for i in range(n):
            carry, y = f(carry, None if xs is None else _tree_index(xs, i))
            ys.append(y)
        return carry, _tree_stack(ys)

    @staticmethod
    def stop_gradient(x):
        return x


lax = _Lax("jax.lax")

tree_util = types.ModuleType("jax.tree_util")
tree_util.Partial = functools.partial

scipy = types.ModuleType("jax.scipy")
scipy.special = types.ModuleType("jax.scipy.special")
def _logsumexp(a, **kw):
    if _isdual(a):
        from ._dual import logsumexp as dlse
        return dlse(a)
    return _wrap(_sps.logsumexp(a, **kw))


scipy.special.logsumexp = _logsumexp

nn = types.ModuleType("jax.nn")
nn.logsumexp = _logsumexp

typing_ = types.ModuleType("jax.typing")
typing_.ArrayLike = typing.Any

random = types.ModuleType("jax.random")


class _Config:
    jax_enable_x64 = True


config = _Config()

import sys as _sys

_sys.modules["jax.numpy"] = numpy
_sys.modules["jax.numpy.linalg"] = _linalg
_sys.modules["jax.lax"] = lax
_sys.modules["jax.tree_util"] = tree_util
_sys.modules["jax.scipy"] = scipy
_sys.modules["jax.scipy.special"] = scipy.special
_sys.modules["jax.nn"] = nn
_sys.modules["jax.typing"] = typing_
_sys.modules["jax.random"] = random
typing = typing_
