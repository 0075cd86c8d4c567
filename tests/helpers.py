"""Shared test helpers: seeded synthetic PPCA data in the reference sampler's shape (ppca_model.rs:164-191)."""
from __future__ import annotations

import numpy as np


def make_data(n, d, k, p_missing=0.2, seed=0, sigma=0.1, mean_scale=0.5, empty_rows=(), empty_dims=(), bern=True):
    rng = np.random.default_rng(seed)
    if bern:
        C_true = (rng.random((d, k)) < 0.3).astype(np.float64) + 0.1 * rng.standard_normal((d, k))
    else:
        C_true = rng.standard_normal((d, k))
    mu_true = mean_scale * rng.standard_normal(d)
    X = rng.standard_normal((n, k)) @ C_true.T + mu_true + sigma * rng.standard_normal((n, d))
    X[rng.random((n, d)) < p_missing] = np.nan
    for r in empty_rows:
        X[r, :] = np.nan
    for c in empty_dims:
        X[:, c] = np.nan
    return X


def init_model(d, k, seed=1000, empty_dims=()):
    """ppca_model.rs:51-70 init semantics with an injected (seeded) C0."""
    rng = np.random.default_rng(seed)
    C0 = rng.standard_normal((d, k))
    for c in empty_dims:
        C0[c, :] = 0.0
    return C0, np.zeros(d), 1.0


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / denom)
