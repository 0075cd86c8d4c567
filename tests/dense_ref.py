"""Independent dense restatement of masked PPCA (numpy, d_obs x d_obs formulas) used to cross-check the oracle.

Nothing here shares code with oracle/ppca_oracle.c: it forms the observed covariance S = sigma^2 I + C_o C_o^T
explicitly and uses slogdet / solve, i.e. the textbook definitions the reference's Woodbury / determinant-lemma
shortcuts (output_covariance.rs:72-121) are derived from.
"""
import numpy as np


def llk_one(x, C, mu, sigma):
    m = np.isfinite(x)
    if not m.any():
        return 0.0
    Co, r = C[m], x[m] - mu[m]
    S = sigma ** 2 * np.eye(m.sum()) + Co @ Co.T
    _, logdet = np.linalg.slogdet(S)
    return -0.5 * r @ np.linalg.solve(S, r) - 0.5 * logdet - 0.5 * np.log(2 * np.pi) * m.sum()


def infer_one(x, C, mu, sigma):
    k = C.shape[1]
    m = np.isfinite(x)
    if not m.any():
        return np.zeros(k), np.eye(k)
    Co, r = C[m], x[m] - mu[m]
    S = sigma ** 2 * np.eye(m.sum()) + Co @ Co.T
    T = Co.T @ np.linalg.inv(S)
    return T @ r, np.eye(k) - T @ Co


def iterate(X, w, C, mu, sigma, tau=0.0, alpha=None, beta=None, m0=None, cov0=None):
    n, d = X.shape
    k = C.shape[1]
    Z = np.zeros((n, k)); COV = np.zeros((n, k, k))
    for i in range(n):
        Z[i], COV[i] = infer_one(X[i], C, mu, sigma)
    M = np.isfinite(X)
    Xc = np.where(M, X - mu, 0.0)
    tcm = (Xc * w[:, None]).T @ Z
    Cn = np.zeros_like(C)
    for i in range(d):
        S = tau * np.eye(k)
        for s in np.nonzero(M[:, i])[0]:
            S += w[s] * (np.outer(Z[s], Z[s]) + COV[s])
        Cn[i] = np.linalg.solve(S, tcm[i]) if np.any(S != 0) else C[i]
    sq = dev2 = 0.0
    tdev = np.zeros(d); tot = np.zeros(d)
    for s in range(n):
        m = M[s]
        if not m.any():
            continue
        Co = C[m]
        sq += w[s] * np.trace(Co @ COV[s] @ Co.T)
        dev = np.where(m, X[s] - C @ Z[s] - mu, 0.0)
        dev2 += w[s] * dev @ dev
        tdev += w[s] * dev
        tot += w[s] * m
    if alpha is not None:
        s2 = ((sq + dev2) / 2 + beta) / (tot.sum() / 2 + alpha + 1)
    else:
        s2 = (sq + dev2) / tot.sum()
    mun = np.where(tot > 0, tdev / np.where(tot > 0, tot, 1), 0.0) + mu
    if m0 is not None:
        P0 = np.linalg.inv(cov0)
        P = np.diag(tot) / s2
        mun = np.linalg.solve(P0 + P, P0 @ m0 + P @ mun)
    return Cn, mun, np.sqrt(s2)
