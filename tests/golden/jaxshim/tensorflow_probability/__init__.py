"""Import-time stand-in for tensorflow_probability (substrates.jax) — test infrastructure.

Only what the reference's `dist.py` / `gam/dist.py` touch on the log-density path:
`tfd.Distribution` (public method → underscore method forwarding) and `tfd.Normal`
(`log_prob = -0.5*((x-loc)/scale)**2 - (0.5*log(2*pi) + log(scale))`, tfp's arrangement,
restated from the published source)."""
import math
import sys
import types

import numpy as np
import scipy.special as sps


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class Distribution:
    def __init__(self, dtype=None, reparameterization_type=None, validate_args=False,
                 allow_nan_stats=True, parameters=None, name=None, **_kw):
        self._dtype = dtype
        self._parameters = parameters
        self._name = name

    @property
    def dtype(self):
        return self._dtype

    @property
    def batch_shape(self):
        return []

    @property
    def event_shape(self):
        return []

    def log_prob(self, value, **kw):
        return self._log_prob(value, **kw)

    def prob(self, value, **kw):
        return self._prob(value, **kw)

    def cdf(self, value, **kw):
        return self._cdf(value, **kw)

    def log_cdf(self, value, **kw):
        return self._log_cdf(value, **kw)

    def quantile(self, value, **kw):
        return self._quantile(value, **kw)

    def mean(self, **kw):
        return self._mean(**kw)

    def stddev(self, **kw):
        return self._stddev(**kw)

    def variance(self, **kw):
        return self._variance(**kw)


class Normal(Distribution):
    def __init__(self, loc, scale, validate_args=False, allow_nan_stats=True, name="Normal"):
        keep = lambda v: v if getattr(v, "_is_dual", False) else np.asarray(v, dtype=np.float64)
        self.loc = keep(loc)
        self.scale = keep(scale)
        super().__init__(dtype=np.float64, name=name)

    def _log_prob(self, x):
        z = (x - self.loc) / self.scale
        return -0.5 * z * z - (0.5 * math.log(2.0 * math.pi) + np.log(self.scale))

    def _cdf(self, x):
        return sps.ndtr((np.asarray(x) - self.loc) / self.scale)

    def _quantile(self, p):
        return sps.ndtri(np.asarray(p)) * self.scale + self.loc

    def _mean(self):
        return self.loc

    def _stddev(self):
        return self.scale

    def _variance(self):
        return self.scale ** 2


FULLY_REPARAMETERIZED = "FULLY_REPARAMETERIZED"


class ParameterProperties:
    def __init__(self, **kw):
        self.kw = kw


class _TensorShape(tuple):
    pass


_tfd = _mod("tensorflow_probability.substrates.jax.distributions",
            Distribution=Distribution, Normal=Normal,
            FULLY_REPARAMETERIZED=FULLY_REPARAMETERIZED)
_tfb = _mod("tensorflow_probability.substrates.jax.bijectors")
_tf = _mod("tensorflow_probability.substrates.jax.tf2jax",
           TensorShape=_TensorShape, reshape=lambda x, s: np.reshape(x, s))
_pp = _mod("tensorflow_probability.substrates.jax.internal.parameter_properties",
           ParameterProperties=ParameterProperties)
_int = _mod("tensorflow_probability.substrates.jax.internal", parameter_properties=_pp)
_sj = _mod("tensorflow_probability.substrates.jax", distributions=_tfd, bijectors=_tfb,
           tf2jax=_tf, internal=_int)
_sub = _mod("tensorflow_probability.substrates", jax=_sj)
_rep = _mod("tensorflow_probability.python.internal.reparameterization",
            FULLY_REPARAMETERIZED=FULLY_REPARAMETERIZED)
_pint = _mod("tensorflow_probability.python.internal", reparameterization=_rep)
_py = _mod("tensorflow_probability.python", internal=_pint)
substrates = _sub
python = _py
