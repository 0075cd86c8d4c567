"""One EM step + one extrapolate per shape-dependent kernel variant, small enough for compute-sanitizer.

    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py

Prints the per-variant launch counters at the end so the log shows which kernels the run covered.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ppca_rs_b200 as pk  # noqa: E402
from helpers import init_model, make_data  # noqa: E402

CASES = [
    # n, d, k, gemm, slices : what it exercises
    (2048, 200, 16, "tc", 6),    # tbitgemm_atm_kernel<6,1,2> (drain), solve_reg16
    (1024, 700, 7, "tc", 6),     # tbitgemm_atm_kernel<6,2,1> (feed: one q tile), solve_reg8
    (2048, 1024, 32, "tc", 6),   # tbitgemm_atm2_kernel (E-step), solve_reg32
    (1024, 260, 64, "tc", 6),    # solve_tile_kernel<64>
    (512, 150, 66, "tc", 6),     # solve_kernel (generic), colmax_kernel_tc; E-step only under the sanitizer (the M-step's
                                 # cross_resid tile at k > 64 asks for more shared memory than the tool leaves available)
    (1024, 200, 16, "tc", 7),    # tbitgemm_kernel<7> (A tile in shared memory)
    (1024, 200, 16, "tc", 8),    # tbitgemm_kernel<8>
    (1024, 200, 16, "int8", 6),  # ibitgemm
    (1024, 200, 16, "dmma", 6),  # bitgemm
]


def main():
    ctx = pk.get_context()
    only = os.environ.get("SANITIZE_ONLY")
    for i, (n, d, k, gemm, slices) in enumerate(CASES):
        if only and str(i) not in only.split(","):
            continue
        ctx.set_gemm(gemm, slices)
        X = make_data(n, d, min(k, 8), 0.25, seed=i, empty_rows=(1,))
        C0, mu0, s0 = init_model(d, k)
        ds = pk.Dataset(X)
        model = pk.PPCAModel(s0, C0, mu0)
        if k > 64 and os.environ.get("SANITIZE_FULL_K", "0") != "1":
            new, llk = model, model.llk(ds)
        else:
            new, llk = model._iterate(ds, None)
        ex = new.extrapolate(ds).numpy()
        assert np.isfinite(ex).all() and np.isfinite(llk)
        print("case", i, (n, d, k, gemm, slices), "llk", llk, flush=True)
    ctx.set_gemm("tc", 6)
    if os.environ.get("SANITIZE_EXTRA", "1") == "1" and (not only or "x" in only.split(",")):
        # one feature in other units: precision ladder (8 planes, DMMA) and the exact-residual pass
        X = make_data(1024, 40, 5, 0.3, seed=3)
        C0, mu0, s0 = init_model(40, 5)
        X[:, 7] *= 1e3
        C0 = C0.copy()
        C0[7] *= 1e3
        new, llk = pk.PPCAModel(1.0, C0, mu0)._iterate(pk.Dataset(X), None)
        print("ill-scaled ok", llk, flush=True)
        # k = 48 (24-column halves of the 128-thread solve), k = 32
        for kk in (48, 32, 40):
            X = make_data(768, 130, 8, 0.25, seed=kk)
            C0, mu0, s0 = init_model(130, kk)
            new, llk = pk.PPCAModel(s0, C0, mu0)._iterate(pk.Dataset(X), None)
            print("k", kk, "ok", llk, flush=True)
        # mean prior through the device Cholesky (3 panels), packed host streaming, one-pass reconstruct with reuse
        rng = np.random.default_rng(1)
        d = 70
        X = make_data(1024, d, 6, 0.25, seed=11)
        C0, mu0, s0 = init_model(d, 6)
        A = rng.standard_normal((d, d))
        prior = pk.Prior().with_mean_prior(rng.standard_normal(d), A @ A.T / d + np.eye(d)).with_isotropic_noise_prior(2.0, 1.0)
        model = pk.PPCAModel(s0, C0, mu0).iterate_with_prior(pk.Dataset(X), prior)
        ctx.set_chunk(512)
        model = model.iterate(pk.HostDataset(X, pin=True, packed=True))
        ctx.set_chunk(0)
        ds = pk.Dataset(X)
        ex, ll = model.reconstruct(ds, True, with_llks=True)
        ex, ll = model.reconstruct(ds, False, out=ex, with_llks=True)
        assert np.isfinite(ex.numpy()).all() and np.isfinite(ll).all()
        print("prior / packed host / reconstruct ok", flush=True)
    if os.environ.get("SANITIZE_MIX", "1") == "1" and not only:
        X = make_data(1500, 64, 4, 0.2, seed=9)
        mix = pk.PPCAMix([pk.PPCAModel(1.0, *init_model(64, kk, seed=j)[:2][::1]) for j, kk in enumerate((4, 6, 3))],
                         np.zeros(3))
        mix2 = mix.iterate(pk.Dataset(X))
        ctx.set_chunk(512)                      # several chunks: running maxima and rescaling
        mix3 = mix2.iterate(pk.Dataset(X))
        ctx.set_chunk(0)
        print("mixture ok", mix3.log_weights, flush=True)
    if os.environ.get("SANITIZE_R2B", "1") == "1" and not only:
        # later round-2 kernels: register-tiled solve (k = 40 / 48 / 64 run above and in CASES), device samplers with the
        # in-kernel Cholesky, full covariances, the fast generator and the regenerated-chunk EM, device ingestion, and the
        # CUDA-graph replay of the mixture chunk loop
        import torch
        rng = np.random.default_rng(4)
        d = 40
        models = [pk.PPCAModel(0.4 + 0.1 * j, rng.standard_normal((d, kk)), rng.standard_normal(d)) for j, kk in enumerate((3, 5))]
        mix = pk.PPCAMix(models, np.log([0.3, 0.7]))
        data = mix.sample(700, 0.3, seed=2)
        infm = mix.infer(data)
        draw = infm.posterior_sampler().sample(seed=3)
        one = models[1].infer(data)
        draw1 = one.posterior_sampler().sample(seed=4)
        cov = one.extrapolated_covariances(models[1], data)
        assert np.isfinite(draw.numpy()).all() and np.isfinite(draw1.numpy()).all() and len(cov) == 700
        print("samplers / full covariances ok", flush=True)
        gen = pk.GeneratedDataset(1500, 70, 6, 0.1, 0.3, seed=5, row_begin=100)
        C0, mu0, s0 = init_model(70, 5)
        ctx.set_chunk(512)
        m2, llk = pk.PPCAModel(s0, C0, mu0)._iterate(gen, None)
        ctx.set_chunk(0)
        t = torch.from_numpy(gen.materialize().numpy()).cuda()
        ds = pk.Dataset.from_device(t[:, 3:60], torch.rand(1500, dtype=torch.float64, device="cuda") + 0.5)
        back = ds.to_torch()
        assert back.shape == (1500, 57) and np.isfinite(llk)
        print("generated / device ingestion ok", flush=True)
        X = make_data(1500, 64, 4, 0.2, seed=9)
        mixg = pk.PPCAMix([pk.PPCAModel(1.0, *init_model(64, kk, seed=j)[:2][::1]) for j, kk in enumerate((4, 6, 3))], np.zeros(3))
        dsg = pk.Dataset(X)
        for _ in range(3):                      # eager, captured, replayed
            mixg2 = mixg.iterate(dsg)
        print("graph replays", ctx.variant_counts()["graph_replays"], flush=True)
    print("variant counts", {k: v for k, v in ctx.variant_counts().items() if v}, flush=True)


if __name__ == "__main__":
    main()
