"""ctypes front end of oracle/libppca_oracle.so.

TEST INFRASTRUCTURE ONLY.  Importers allowed: tests/, __graft_entry__.smoke(), and bench.py's
cpu_baseline / --impl reference legs.  The product package (ppca_rs_b200) must never import this.

Every function restates a reference symbol; the citation is in ppca_oracle.c next to the C body.
"parity unpinned" beyond the two KATs of ppca/src/ppca_model.rs:658-671 (see the C header).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libppca_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "ppca_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "clean", "all"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class _Prior(C.Structure):
    _fields_ = [
        ("has_mean_prior", C.c_int),
        ("mean", _dp),
        ("mean_prec", _dp),
        ("has_noise_prior", C.c_int),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("transformation_precision", C.c_double),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_quadratic_form.restype = C.c_double
        _lib.oracle_covariance_log_det.restype = C.c_double
        _lib.oracle_llk.restype = C.c_double
        _lib.oracle_mix_llk.restype = C.c_double
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_dp)


@dataclass
class Prior:
    """prior.rs:8-29."""

    mean: Optional[np.ndarray] = None
    mean_covariance: Optional[np.ndarray] = None
    isotropic_noise_alpha: Optional[float] = None
    isotropic_noise_beta: Optional[float] = None
    transformation_precision: float = 0.0

    def _c(self):
        keep = []
        pr = _Prior()
        pr.has_mean_prior = 0
        if self.mean is not None:
            m = _f64(self.mean).reshape(-1)
            prec = _f64(np.linalg.inv(_f64(self.mean_covariance)))  # prior.rs:36-41 try_inverse
            keep += [m, prec]
            pr.has_mean_prior = 1
            pr.mean = _p(m)
            pr.mean_prec = _p(prec)
        pr.has_noise_prior = int(self.isotropic_noise_alpha is not None)
        pr.alpha = float(self.isotropic_noise_alpha or 0.0)
        pr.beta = float(self.isotropic_noise_beta or 0.0)
        pr.transformation_precision = float(self.transformation_precision)
        return pr, keep


class stable:
    """Context manager: run the oracle with the cancellation-free posterior formulas (see ppca_oracle.c)."""

    def __enter__(self):
        lib().oracle_set_stable(1)
        return self

    def __exit__(self, *exc):
        lib().oracle_set_stable(0)
        return False


def set_num_threads(n: int) -> None:
    """rayon's global pool uses every core regardless of OMP_NUM_THREADS (which torchrun sets to 1)."""
    lib().oracle_set_num_threads.argtypes = [C.c_int]
    lib().oracle_set_num_threads.restype = None
    lib().oracle_set_num_threads(int(n))


def num_threads() -> int:
    return lib().oracle_num_threads()


def quadratic_form(Cm, sigma, x) -> float:
    Cm = _f64(Cm); x = _f64(x)
    d, k = Cm.shape
    return lib().oracle_quadratic_form(d, k, _p(Cm), C.c_double(sigma), _p(x))


def covariance_log_det(Cm, sigma) -> float:
    Cm = _f64(Cm)
    d, k = Cm.shape
    return lib().oracle_covariance_log_det(d, k, _p(Cm), C.c_double(sigma))


def empty_dimensions(X) -> list:
    X = _f64(X)
    n, d = X.shape
    out = np.zeros(d, dtype=np.uint8)
    lib().oracle_empty_dimensions(C.c_int64(n), d, _p(X), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return [int(i) for i in np.nonzero(out)[0]]


def llks(X, Cm, mu, sigma) -> np.ndarray:
    X = _f64(X); Cm = _f64(Cm); mu = _f64(mu).reshape(-1)
    n, d = X.shape
    out = np.empty(n)
    lib().oracle_llks(C.c_int64(n), d, Cm.shape[1], _p(X), _p(Cm), _p(mu), C.c_double(sigma), _p(out))
    return out


def llk(X, w, Cm, mu, sigma) -> float:
    X = _f64(X); Cm = _f64(Cm); mu = _f64(mu).reshape(-1)
    n, d = X.shape
    w = np.ones(n) if w is None else _f64(w)
    return lib().oracle_llk(C.c_int64(n), d, Cm.shape[1], _p(X), _p(w), _p(Cm), _p(mu), C.c_double(sigma))


def infer(X, Cm, mu, sigma):
    X = _f64(X); Cm = _f64(Cm); mu = _f64(mu).reshape(-1)
    n, d = X.shape
    k = Cm.shape[1]
    Z = np.empty((n, k)); COV = np.empty((n, k, k))
    lib().oracle_infer(C.c_int64(n), d, k, _p(X), _p(Cm), _p(mu), C.c_double(sigma), _p(Z), _p(COV))
    return Z, COV


def smooth(X, Cm, mu, sigma, extrapolate=False) -> np.ndarray:
    X = _f64(X); Cm = _f64(Cm); mu = _f64(mu).reshape(-1)
    n, d = X.shape
    out = np.empty((n, d))
    lib().oracle_smooth(C.c_int64(n), d, Cm.shape[1], _p(X), _p(Cm), _p(mu), C.c_double(sigma),
                        int(extrapolate), _p(out))
    return out


def extrapolate(X, Cm, mu, sigma) -> np.ndarray:
    return smooth(X, Cm, mu, sigma, extrapolate=True)


def iterate(X, w, Cm, mu, sigma, prior: Optional[Prior] = None):
    """Returns (C', mu', sigma')."""
    X = _f64(X); Cm = _f64(Cm); mu = _f64(mu).reshape(-1)
    n, d = X.shape
    k = Cm.shape[1]
    w = np.ones(n) if w is None else _f64(w)
    C_out = np.empty((d, k)); mu_out = np.empty(d); s_out = C.c_double(0.0)
    pr_ref = None
    keep = None
    if prior is not None:
        pr, keep = prior._c()
        pr_ref = C.byref(pr)
    rc = lib().oracle_iterate(C.c_int64(n), d, k, _p(X), _p(w), _p(Cm), _p(mu), C.c_double(sigma), pr_ref,
                              _p(C_out), _p(mu_out), C.byref(s_out))
    if rc:
        raise RuntimeError(f"oracle_iterate failed rc={rc}")
    return C_out, mu_out, s_out.value


def to_canonical(Cm) -> np.ndarray:
    Cm = _f64(Cm)
    d, k = Cm.shape
    out = np.empty((d, k))
    lib().oracle_to_canonical(d, k, _p(Cm), _p(out))
    return out


def log_softmax(v) -> np.ndarray:
    v = _f64(v).copy()
    lib().oracle_log_softmax(_p(v), len(v))
    return v


# ---- mixtures: models = list of (C, mu, sigma); logw normalised ----

def _pack(models: Sequence):
    ks = np.array([np.asarray(m[0]).shape[1] for m in models], dtype=np.int32)
    Cs = np.concatenate([_f64(m[0]).reshape(-1) for m in models]) if len(models) else np.zeros(0)
    mus = _f64(np.stack([_f64(m[1]).reshape(-1) for m in models]))
    sig = _f64([m[2] for m in models])
    return ks, _f64(Cs), mus, sig


def mix_llks(X, models, logw) -> np.ndarray:
    X = _f64(X); n, d = X.shape
    ks, Cs, mus, sig = _pack(models); logw = _f64(logw)
    out = np.empty(n)
    lib().oracle_mix_llks(C.c_int64(n), d, len(models), ks.ctypes.data_as(_ip), _p(X), _p(Cs), _p(mus),
                          _p(sig), _p(logw), _p(out))
    return out


def mix_llk(X, w, models, logw) -> float:
    X = _f64(X); n, d = X.shape
    w = np.ones(n) if w is None else _f64(w)
    ks, Cs, mus, sig = _pack(models); logw = _f64(logw)
    return lib().oracle_mix_llk(C.c_int64(n), d, len(models), ks.ctypes.data_as(_ip), _p(X), _p(w), _p(Cs),
                                _p(mus), _p(sig), _p(logw))


def mix_infer_cluster(X, models, logw) -> np.ndarray:
    X = _f64(X); n, d = X.shape
    ks, Cs, mus, sig = _pack(models); logw = _f64(logw)
    out = np.empty((n, len(models)))
    lib().oracle_mix_infer_cluster(C.c_int64(n), d, len(models), ks.ctypes.data_as(_ip), _p(X), _p(Cs),
                                   _p(mus), _p(sig), _p(logw), _p(out))
    return out


def mix_smooth(X, models, logw, extrapolate=False) -> np.ndarray:
    X = _f64(X); n, d = X.shape
    ks, Cs, mus, sig = _pack(models); logw = _f64(logw)
    out = np.empty((n, d))
    lib().oracle_mix_smooth(C.c_int64(n), d, len(models), ks.ctypes.data_as(_ip), _p(X), _p(Cs), _p(mus),
                            _p(sig), _p(logw), int(extrapolate), _p(out))
    return out


def mix_iterate(X, w, models, logw, prior: Optional[Prior] = None):
    """Returns (models', logw')."""
    X = _f64(X); n, d = X.shape
    w = np.ones(n) if w is None else _f64(w)
    ks, Cs, mus, sig = _pack(models); logw = _f64(logw)
    Cs_o = np.empty_like(Cs); mus_o = np.empty_like(mus); sig_o = np.empty_like(sig); lw_o = np.empty_like(logw)
    pr_ref = None
    keep = None
    if prior is not None:
        pr, keep = prior._c()
        pr_ref = C.byref(pr)
    rc = lib().oracle_mix_iterate(C.c_int64(n), d, len(models), ks.ctypes.data_as(_ip), _p(X), _p(w), _p(Cs),
                                  _p(mus), _p(sig), _p(logw), pr_ref, _p(Cs_o), _p(mus_o), _p(sig_o), _p(lw_o))
    if rc:
        raise RuntimeError(f"oracle_mix_iterate failed rc={rc}")
    out = []
    off = 0
    for j, k in enumerate(ks):
        out.append((Cs_o[off:off + d * k].reshape(d, k).copy(), mus_o[j].copy(), float(sig_o[j])))
        off += d * k
    return out, lw_o


# ---- per-sample output covariances, restated in numpy in the reference's operation order -----------------
def smoothed_covariance(Cm, sigma, cov) -> np.ndarray:
    """InferredMasked::smoothed_covariance (ppca_model.rs:471-477): I sigma^2 + C Sigma C^T."""
    Cm, cov = _f64(Cm), _f64(cov)
    d = Cm.shape[0]
    return np.eye(d) * sigma ** 2 + Cm @ cov @ Cm.T


def smoothed_covariance_diagonal(Cm, sigma, cov) -> np.ndarray:
    """ppca_model.rs:485-508: row-wise dot of (C Sigma) with C, plus sigma^2."""
    Cm, cov = _f64(Cm), _f64(cov)
    return np.einsum("ia,ia->i", Cm @ cov, Cm) + sigma ** 2


def extrapolated_covariance(Cm, sigma, cov, x) -> np.ndarray:
    """ppca_model.rs:517-534: the smoothed covariance of the MISSING dimensions expanded back to d x d (zeros on the rows and
    columns of observed dimensions; all zeros when nothing is missing)."""
    Cm, cov, x = _f64(Cm), _f64(cov), np.asarray(x, dtype=np.float64)
    d = Cm.shape[0]
    neg = ~np.isfinite(x)
    out = np.zeros((d, d))
    if not neg.any():
        return out
    sub = Cm[neg]
    out[np.ix_(neg, neg)] = np.eye(sub.shape[0]) * sigma ** 2 + sub @ cov @ sub.T
    return out


def extrapolated_covariance_diagonal(Cm, sigma, cov, x) -> np.ndarray:
    """ppca_model.rs:542-577."""
    Cm, cov, x = _f64(Cm), _f64(cov), np.asarray(x, dtype=np.float64)
    neg = ~np.isfinite(x)
    out = np.zeros(Cm.shape[0])
    if neg.any():
        sub = Cm[neg]
        out[neg] = np.einsum("ia,ia->i", sub @ cov, sub) + sigma ** 2
    return out


def mix_covariance(post, means, covs) -> np.ndarray:
    """InferredMaskedMix::smoothed_covariance / extrapolated_covariance (mix.rs:422-437, 466-481) for ONE sample:
    sum_j p_j (cov_j + (m_j - mean)(m_j - mean)^T), mean = sum_j p_j m_j; `covs` are the components' SMOOTHED covariances in
    both forms (mix.rs:474), `means` their smoothed (or extrapolated) outputs."""
    post = np.asarray(post, dtype=np.float64)
    mean = sum(p * m for p, m in zip(post, means))
    return sum(p * (c + np.outer(m - mean, m - mean)) for p, m, c in zip(post, means, covs))
