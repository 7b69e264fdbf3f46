"""Seeded synthetic workloads of SURVEY.md section 8(d) / BASELINE.json `configs`.

Pure NumPy + the host-side set-up code of the package (no oracle import): used by
bench.py's own arm, by the reference arm (which hands the SAME arrays to the
oracle), and by the full-size property tests.
"""

from __future__ import annotations

import dataclasses

import numpy as np

import liesel_ptm_b200 as lp

CONFIGS = {
    # name: (n, n_loc, n_scale, nbases, chains_per_gpu)
    "c1_readme_n100": (100, 1, 0, 20, 4),
    "c2_n100k_3loc_1scale_64chains": (100_000, 3, 1, 20, 64),
    "c3_n1M_5loc_5scale_nb20_1024chains": (1_000_000, 5, 5, 20, 1024),
    "c4_n1M_10loc_1scale_nb30_256chains": (1_000_000, 10, 1, 30, 256),
    # c5 (n = 50M over 8 GPUs, 20 smooth terms, 128 chains): the per-GPU slice is 6.25M
    # observations; built per rank from its own seeded slice (see make_c5_slice)
    "c5_n50M_10loc_10scale_128chains": (50_000_000, 10, 10, 20, 128),
}


def _f(k, x):
    """util/simfun.py:9-32 (f1_linear, f2_ushaped, f3_oscillating, f4_bell), cycled."""
    k %= 4
    if k == 0:
        return 1.0 * x
    if k == 1:
        return x + ((2 * x) ** 2) / 5.5
    if k == 2:
        return -x + np.pi * np.sin(np.pi * x)
    pdf = lambda v: np.exp(-0.5 * v * v) / np.sqrt(2 * np.pi)  # noqa: E731
    return 0.5 * x + 15 * pdf(2 * (x - 0.2)) - pdf(x + 0.4)


def make_data(n, n_loc, n_scale, seed=0, dtype=np.float64):
    """x_t ~ U(-2, 2) iid, bimodal residual mixture (README.md:42-47), response mapped so
    that its 1 % / 99 % quantiles sit at -3.3 / +3.3 (=> ~98 % of the standardised
    residuals in the core [-4, 4], the rest in both transitions and both tails)."""
    rng = np.random.default_rng(seed)
    T = n_loc + n_scale
    X = rng.uniform(-2.0, 2.0, size=(T, n))
    mu = np.zeros(n)
    for t in range(n_loc):
        mu += 0.25 * _f(t, X[t]) / max(n_loc, 1)
    ls = np.zeros(n)
    for t in range(n_scale):
        ls += 0.3 * np.sin(X[n_loc + t]) / max(n_scale, 1)
    comp = rng.uniform(size=n) < 0.5
    res = np.where(comp, rng.normal(2.0, 1.0, size=n), rng.normal(-2.0, 0.5, size=n))
    y = mu + np.exp(ls) * res
    q01, q99 = np.quantile(y, [0.01, 0.99])
    y = -3.3 + 6.6 * (y - q01) / (q99 - q01)
    return y.astype(dtype), X.astype(dtype)


@dataclasses.dataclass
class Workload:
    name: str
    model: lp.LocScalePTM
    params: np.ndarray  # [C, P]
    y: np.ndarray
    X: np.ndarray
    n: int
    n_loc: int
    n_scale: int
    nbases: int
    n_covariates: int

    @property
    def bytes_alg(self) -> int:
        """Algorithmic bytes per obs-chain evaluation: w * (1 + T_x), SURVEY.md 8(d)."""
        return self.params.dtype.itemsize * (1 + self.n_covariates)


def make_states(model: lp.LocScalePTM, C: int, seed=1, amp=0.1, sample=20000):
    """Seeded chain states: every smooth f_t has sd ~ amp (beta scaled by the design's
    column RMS, estimated on a subsample), delta_z ~ N(0, 0.2^2), tau_t = 1,
    tau_delta = 0.5, intercepts 0."""
    rng = np.random.default_rng(seed)
    betas = []
    for t, _ in model.all_terms():
        b = t.basis
        idx = np.arange(min(sample, b.x.shape[0]))
        j, vals = lp.banded_basis(b.x[idx].astype(np.float64), b.knots.astype(np.float64))
        D = np.zeros((idx.shape[0], b.nbases + 1))
        for m in range(4):
            D[idx, j - 3 + m] = vals[:, m]
        rms = np.sqrt(np.mean((D[:, : b.nbases] @ b.Z.astype(np.float64)) ** 2, axis=0))
        betas.append(amp * rng.normal(size=(C, b.ncoef)) / (rms * np.sqrt(b.ncoef)))
    dz = rng.normal(0, 0.2, size=(C, model.nparam - 1))
    return model.pack(betas, dz, tau=1.0, tau_delta=0.5)


def build(name: str, chains: int | None = None, dtype=np.float64, device="cuda", seed=0,
          n_override: int | None = None) -> Workload:
    n, n_loc, n_scale, nbases, C = CONFIGS[name]
    if n_override:
        n = n_override
    if chains:
        C = chains
    y, X = make_data(n, n_loc, n_scale, seed=seed, dtype=dtype)
    model = lp.LocScalePTM.new_ptm(y, nparam=20, to_float32=(np.dtype(dtype) == np.float32),
                                   device=device)
    for t in range(n_loc + n_scale):
        tm = lp.term.f_ig(lp.ps(X[t], nbases=nbases, xname=f"x{t}", dtype=dtype),
                          fname="f" if t < n_loc else "g")
        if t < n_loc:
            model.loc += tm
        else:
            model.scale += tm
    params = make_states(model, C, seed=seed + 1).astype(dtype)
    return Workload(name=name, model=model, params=params, y=y, X=X, n=n, n_loc=n_loc,
                    n_scale=n_scale, nbases=nbases, n_covariates=n_loc + n_scale)
