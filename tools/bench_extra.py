#!/usr/bin/env python
"""Side measurements for BASELINE configs 4 (PPCAMix) and 5 (inference): not the bench.py contract, just timings."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import ppca_rs_b200 as pk

ctx = pk.get_context()
rng = np.random.default_rng(0)


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


out = {}
# config 5 shape: d=1024, k=48, inference only
n, d, k = 1_000_000, 1024, 48
ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.3, seed=7)
model = pk.PPCAModel(0.5, rng.standard_normal((d, k)), np.zeros(d))
for name, fn in (("llk", lambda: model.llk(ds)), ("extrapolate", lambda: model.extrapolate(ds)), ("smooth", lambda: model.smooth(ds))):
    t = timeit(fn)
    out[f"c5_{name}_samples_per_s"] = n / t
del ds
# config 4 shape: M components, d=512, k=32
for M in (4, 32):
    n = 400_000 if M == 4 else 100_000
    ds = pk.Dataset.synthetic(n, 512, 32, 0.1, 0.25, n_components=M, seed=9)
    mix = pk.PPCAMix([pk.PPCAModel(1.0, rng.standard_normal((512, 32)), 0.1 * rng.standard_normal(512)) for _ in range(M)], np.zeros(M))
    t = timeit(lambda: mix._iterate(ds, None), reps=2)
    out[f"c4_M{M}_samples_iters_per_s"] = n / t
    out[f"c4_M{M}_component_samples_iters_per_s"] = n * M / t
    del ds
print(json.dumps(out))
