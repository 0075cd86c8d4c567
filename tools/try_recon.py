import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ppca_rs_b200 as pk
rng = np.random.default_rng(0)
n, d, k = 300_000, 1024, 48
ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.3, seed=7)
model = pk.PPCAModel(0.5, rng.standard_normal((d, k)), np.zeros(d))
for name, fn in (("llk", lambda: model.llk(ds)), ("extrapolate", lambda: model.extrapolate(ds)), ("smooth", lambda: model.smooth(ds))):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0); del r
    print(name, ["%.1f ms" % (1e3 * t) for t in ts], "best samples/s %.3g" % (n / min(ts)))
