"""CPU stand-in for ppca_rs_b200.distributed.CudaEngine, for the world_size-2 gloo tests of the host logic.

Produces the SAME statistics buffer layout as the CUDA engine (ppca_b200_em_stats_len) with plain numpy
(cancellation-free posterior formulas), and finishes the M-step from a reduced buffer.  Test infrastructure only.
"""
import numpy as np
import torch

from ppca_rs_b200.distributed import stats_len
from ppca_rs_b200.model import PPCAModel

LN_2PI = 1.8378770664093453


class HostShard:
    """Stands in for a device Dataset: a host matrix with NaN = missing, plus weights."""

    def __init__(self, X, w=None):
        self.X = np.asarray(X, dtype=np.float64)
        self.w = np.ones(self.X.shape[0]) if w is None else np.asarray(w, dtype=np.float64)

    def __len__(self):
        return self.X.shape[0]


def _layout(d, k):
    kk = k * (k + 1) // 2
    kkp = (max(kk, 1) + 7) // 8 * 8
    kp = (max(k, 1) + 7) // 8 * 8
    offA, offB = 0, d * kkp
    offT = offB + d * kp
    offO = offT + d
    offS = offO + d
    return kk, kkp, kp, offA, offB, offT, offO, offS


def _posterior(x, C, mu, sigma):
    m = np.isfinite(x)
    k = C.shape[1]
    if not m.any():
        return m, np.zeros(k), np.eye(k), 0.0
    Co, r = C[m], x[m] - mu[m]
    M = sigma ** 2 * np.eye(k) + Co.T @ Co
    Minv = np.linalg.inv(M)
    y = Co.T @ r
    z = Minv @ y
    _, logdet = np.linalg.slogdet(M)
    dn = int(m.sum())
    llk = -0.5 * (r @ r - y @ z) / sigma ** 2 - 0.5 * (logdet + 2 * np.log(sigma) * (dn - k)) - 0.5 * LN_2PI * dn
    return m, z, sigma ** 2 * Minv, llk


class NumpyEngine:
    def new_stats(self, d, k):
        return torch.zeros(stats_len(d, k), dtype=torch.float64)

    def em_stats(self, ds, model, stats):
        C, mu, sigma = model._C, model._mu, model._sigma
        d, k = C.shape
        kk, kkp, kp, offA, offB, offT, offO, offS = _layout(d, k)
        buf = np.zeros(stats_len(d, k))
        A = buf[offA:offB].reshape(d, kkp)
        B = buf[offB:offT].reshape(d, kp)
        iu = np.triu_indices(k)
        for x, w in zip(ds.X, ds.w):
            m, z, cov, llk = _posterior(x, C, mu, sigma)
            buf[offS + 2] += w * llk
            buf[offS + 3] += w
            if not m.any():
                continue
            buf[offS + 4] += 1.0
            W = w * (np.outer(z, z) + cov)
            A[m, :kk] += W[iu]
            xc = np.where(m, x - mu, 0.0)
            B[:, :k] += w * np.outer(xc, z)
            Co = C[m]
            buf[offS + 0] += w * np.trace(Co @ cov @ Co.T)
            dev = np.where(m, x - C @ z - mu, 0.0)
            buf[offS + 1] += w * dev @ dev
            buf[offT:offO] += w * dev
            buf[offO:offS] += w * m
        stats.copy_(torch.from_numpy(buf))

    def em_finish(self, model, prior, stats):
        assert prior is None
        C, mu = model._C, model._mu
        d, k = C.shape
        kk, kkp, kp, offA, offB, offT, offO, offS = _layout(d, k)
        buf = stats.numpy()
        A = buf[offA:offB].reshape(d, kkp)
        B = buf[offB:offT].reshape(d, kp)
        iu = np.triu_indices(k)
        Cn = np.empty_like(C)
        for i in range(d):
            S = np.zeros((k, k))
            S[iu] = A[i, :kk]
            S = S + S.T - np.diag(np.diag(S))
            Cn[i] = np.linalg.solve(S, B[i, :k]) if np.any(S != 0) else C[i]
        tdev, tot, sc = buf[offT:offO], buf[offO:offS], buf[offS:]
        s2 = (sc[0] + sc[1]) / tot.sum()
        mun = np.where(tot > 0, tdev / np.where(tot > 0, tot, 1.0), 0.0) + mu
        return PPCAModel(float(np.sqrt(s2)), Cn, mun), float(sc[2])
