import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ppca_rs_b200 as pk
ctx = pk.get_context()
def rel(a,b): return float(np.max(np.abs(np.asarray(a)-np.asarray(b)))/np.max(np.abs(b)))
for (n,d,k) in [(1000,40,4),(3000,200,16),(2000,300,32),(1500,260,64)]:
    ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.25, seed=3)
    rng = np.random.default_rng(1)
    model = pk.PPCAModel(0.7, rng.standard_normal((d,k)), 0.1*rng.standard_normal(d))
    ctx.set_gemm("dmma"); a, la = model._iterate(ds, None); lla = model.llks(ds)
    for T in (6, 7, 8):
        ctx.set_gemm("tc", T); b, lb = model._iterate(ds, None); llb = model.llks(ds)
        print(n,d,k,"T",T,"C",rel(b.transform,a.transform),"mu",rel(b.mean,a.mean),"s",abs(b.isotropic_noise-a.isotropic_noise)/a.isotropic_noise,"llk",abs(lb-la)/abs(la),"llks",rel(llb,lla), flush=True)
ctx.set_gemm("dmma")
print("done")
