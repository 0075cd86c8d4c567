#!/usr/bin/env python
"""Generates tests/golden/*.npz with the CPU oracle (oracle/ppca_oracle.c, faithful operation order).

The Rust reference cannot be imported or built in this image, so these vectors are the ORACLE's outputs on
seeded inputs, not the reference's: they pin the oracle against drift and give the GPU tests fixed targets
("parity unpinned" beyond the two reference KATs, see the oracle header).  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from helpers import init_model, make_data  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def single(name, n, d, k, p, iters, **kw):
    X = make_data(n, d, k, p, seed=7, **kw)
    C, mu, s = init_model(d, k, empty_dims=kw.get("empty_dims", ()))
    w = np.random.default_rng(3).random(n) + 0.5
    out = dict(X=X, w=w, C0=C, mu0=mu, s0=np.array(s))
    out["llks0"] = orc.llks(X, C, mu, s)
    out["Z0"], out["COV0"] = orc.infer(X, C, mu, s)
    out["smooth0"] = orc.smooth(X, C, mu, s)
    out["extrapolate0"] = orc.extrapolate(X, C, mu, s)
    Cs, mus, ss, llk = [], [], [], []
    for _ in range(iters):
        llk.append(orc.llk(X, w, C, mu, s))
        C, mu, s = orc.iterate(X, w, C, mu, s)
        Cs.append(C); mus.append(mu); ss.append(s)
    out.update(C_traj=np.stack(Cs), mu_traj=np.stack(mus), s_traj=np.array(ss), llk_traj=np.array(llk))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


def mixture(name, n, d, ks, iters):
    X = np.concatenate([make_data(n // len(ks), d, k, 0.2, seed=20 + j, mean_scale=2.0) for j, k in enumerate(ks)])
    np.random.default_rng(0).shuffle(X, axis=0)
    w = np.random.default_rng(5).random(X.shape[0]) + 0.5
    models = []
    for j, k in enumerate(ks):
        C0, mu0, _ = init_model(d, k, seed=1000 + j)
        models.append((C0, mu0 + 0.1 * j, 1.0 + 0.1 * j))
    logw = orc.log_softmax(np.log(np.arange(1, len(ks) + 1.0)))
    out = dict(X=X, w=w, ks=np.array(ks), logw0=logw)
    for j, (C, mu, s) in enumerate(models):
        out[f"C0_{j}"], out[f"mu0_{j}"], out[f"s0_{j}"] = C, mu, np.array(s)
    out["llks0"] = orc.mix_llks(X, models, logw)
    out["logpost0"] = orc.mix_infer_cluster(X, models, logw)
    out["smooth0"] = orc.mix_smooth(X, models, logw)
    out["extrapolate0"] = orc.mix_smooth(X, models, logw, extrapolate=True)
    for it in range(iters):
        out[f"llk_{it}"] = np.array(orc.mix_llk(X, w, models, logw))
        models, logw = orc.mix_iterate(X, w, models, logw)
        out[f"logw_{it}"] = logw
        for j, (C, mu, s) in enumerate(models):
            out[f"C_{it}_{j}"], out[f"mu_{it}_{j}"], out[f"s_{it}_{j}"] = C, mu, np.array(s)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    single("toy_d3_k2", 100, 3, 2, 0.2, 10)                                   # BASELINE configs[0]
    single("ragged_d37_k5", 300, 37, 5, 0.3, 6, empty_rows=(4,), empty_dims=(36,))
    single("c2shape_d200_k16", 150, 200, 16, 0.2, 3)                          # BASELINE configs[1] shape
    mixture("mix_d12_k232", 300, 12, (2, 3, 2), 3)
    print("golden vectors written to", HERE)
