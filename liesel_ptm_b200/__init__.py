"""liesel_ptm_b200 -- B200-native (sm_100a) hot path of liesel-ptm.

Only what the path needs: the CUDA kernels + C ABI under ``csrc/`` and the
host-side mirror of the reference's operator interface (``LocScalePTM``, ``ps``,
``term``, ``FlatLogProb``, ``GaussianLocCholInfo``).  Importing the package does
not load the CUDA library; ``LocScalePTM.build()`` does and fails loudly if it is
missing (there is no CPU fallback).
"""

from .spline import (  # noqa: F401
    Basis,
    LinearConstraintEVD,
    LogIncKnots,
    Penalty,
    OnionKnots,
    PTMKnots,
    approx_grid,
    banded_basis,
    equidistant_knots,
    lin,
    mixed_model,
    ps,
    term,
    trafo_reparam,
)
from ._lib import PtmError  # noqa: F401
from .model import (  # noqa: F401
    FlatLogProb,
    GaussianLocCholInfo,
    GaussianScaleCholInfo,
    LocScalePTM,
)

__all__ = [
    "LocScalePTM", "FlatLogProb", "GaussianLocCholInfo", "GaussianScaleCholInfo",
    "ps", "lin", "term", "PTMKnots", "OnionKnots", "LogIncKnots", "Penalty", "equidistant_knots",
    "mixed_model", "LinearConstraintEVD", "banded_basis", "approx_grid", "trafo_reparam", "Basis",
    "PtmError",
]
