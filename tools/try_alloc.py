"""Times the pieces of the mixture e2e step (dataset upload from pinned memory, sharded mixture step)."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ppca_rs_b200 as pk
from ppca_rs_b200 import distributed as pdist

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
ctx = pk.Context(0, stream.cuda_stream); pk.set_context(ctx)
n, d, k, m = 200_000, 512, 32, 4
ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.25, n_components=m, seed=1, ctx=ctx)
Xh = ds.numpy()
host = pk.HostDataset(Xh, pin=True, ctx=ctx)
rng = np.random.default_rng(0)
mix = pk.PPCAMix([pk.PPCAModel(1.0, rng.standard_normal((d, k)), np.zeros(d)) for _ in range(m)], np.zeros(m))
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    dsh = pk.Dataset(Xh, _ctx=ctx)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    st = pdist.ShardedPPCAMix(ctx, dsh, mix, group=None)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st.step()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    mix = st.mix
    del dsh, st
    torch.cuda.synchronize(); t4 = time.perf_counter()
    print(f"it {it}: upload {1e3*(t1-t0):.1f} ms, construct {1e3*(t2-t1):.1f} ms, step {1e3*(t3-t2):.1f} ms, free {1e3*(t4-t3):.1f} ms", flush=True)
