"""GPU parity: the CUDA engine (through the C ABI / ppca_rs_b200 front end) against the CPU oracle.

Tolerance: BASELINE.json north_star — 1e-9 relative for the FP64 path (per-iteration llk, C, mu, sigma^2),
identical masks and bit-identical observed slots for extrapolate.
"""
import numpy as np
import pytest

from helpers import init_model, make_data, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9


def both(orc, fn, *a, **kw):
    """(faithful, stable) oracle results.  `faithful` follows the reference's operation order, whose own
    cancellation noise (I - T C_o, output_covariance.rs:90-101) grows like eps (|G|/sigma^2)^2 and passes 1e-9
    once sigma is small; `stable` is the algebraically identical cancellation-free form (see ppca_oracle.c)."""
    faithful = fn(*a, **kw)
    with orc.stable():
        st = fn(*a, **kw)
    return faithful, st


def assert_close(got, faithful, stable, what=""):
    """1e-9 against the stable oracle; 1e-9 plus the reference's own noise floor against the faithful one."""
    scale = max(np.max(np.abs(np.asarray(stable))), 1e-300)
    gap = np.max(np.abs(np.asarray(faithful) - np.asarray(stable))) / scale
    err_s = np.max(np.abs(np.asarray(got) - np.asarray(stable))) / scale
    err_f = np.max(np.abs(np.asarray(got) - np.asarray(faithful))) / scale
    assert err_s < TOL, f"{what}: {err_s:.3e} vs stable oracle"
    assert err_f < TOL + 2.0 * gap, f"{what}: {err_f:.3e} vs faithful oracle (its noise floor {gap:.3e})"


SHAPES = [
    # n, d, k, p_missing
    (100, 3, 2, 0.2),      # config 1: examples/toy_model.py
    (777, 37, 5, 0.3),     # ragged everything
    (3000, 200, 16, 0.2),  # config 2 shape, small n
    (1500, 70, 10, 0.25),
    (1200, 150, 32, 0.25), # config 4 state size
    (600, 130, 48, 0.3),   # config 5 state size
    (500, 260, 64, 0.3),   # config 3 state size
    (300, 20, 1, 0.1),
]


@pytest.fixture(scope="module")
def pk():
    import ppca_rs_b200 as pk
    return pk


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _case(n, d, k, p, seed=0, **kw):
    X = make_data(n, d, k, p, seed=seed, **kw)
    C0, mu0, s0 = init_model(d, k, empty_dims=kw.get("empty_dims", ()))
    return X, C0, mu0, s0


def test_dataset_roundtrip(pk, orc):
    X = make_data(1000, 45, 4, 0.3, seed=3, empty_rows=(5, 17), empty_dims=(7, 44))
    X[3, 2] = np.inf
    X[4, 1] = -np.inf
    w = np.random.default_rng(1).random(1000) + 0.5
    ds = pk.Dataset(X, w)
    assert len(ds) == 1000 and ds.output_size() == 45
    back = ds.numpy()
    fin = np.isfinite(X)
    assert np.array_equal(np.isfinite(back), fin)            # identical masks
    assert np.array_equal(back[fin], X[fin])                  # bit-identical observed values
    assert np.isnan(back[~fin]).all()                         # masked_vector: NaN (also where the input was inf)
    assert np.array_equal(ds.weights(), w)
    assert ds.empty_dimensions() == orc.empty_dimensions(X) == [7, 44]
    # chunks / concat (src/python_bindings.rs:110-133,151-165)
    parts = list(ds.chunks(3))
    assert [len(p) for p in parts] == [334, 334, 332]
    cat = pk.Dataset.concat(parts)
    assert np.array_equal(np.isnan(cat.numpy()), np.isnan(back))
    assert np.array_equal(cat.numpy()[fin], X[fin])
    assert np.array_equal(cat.weights(), w)
    # bincode round trip
    again = pk.Dataset.load(ds.dump())
    assert np.array_equal(again.numpy()[fin], X[fin]) and np.array_equal(again.weights(), w)


@pytest.mark.parametrize("n,d,k,p", SHAPES)
def test_llks_and_llk(pk, orc, n, d, k, p):
    X, C0, mu0, s0 = _case(n, d, k, p, empty_rows=(1,))
    w = np.random.default_rng(7).random(n) + 0.25
    ds = pk.Dataset(X, w)
    for sigma in (1.0, 0.3):
        model = pk.PPCAModel(sigma, C0, mu0.reshape(1, -1))
        got = model.llks(ds)
        want = orc.llks(X, C0, mu0, sigma)
        assert got[1] == 0.0                                   # empty sample (ppca_model.rs:125-129)
        assert rel_err(got, want) < TOL
        assert abs(model.llk(ds) - orc.llk(X, w, C0, mu0, sigma)) <= TOL * abs(orc.llk(X, w, C0, mu0, sigma))


@pytest.mark.parametrize("n,d,k,p", SHAPES)
def test_infer(pk, orc, n, d, k, p):
    X, C0, mu0, s0 = _case(n, d, k, p, empty_rows=(0, 9))
    ds = pk.Dataset(X)
    model = pk.PPCAModel(0.5, C0, mu0.reshape(-1, 1))
    inf = model.infer(ds)
    (Z, COV), (Zs, COVs) = both(orc, orc.infer, X, C0, mu0, 0.5)
    assert_close(inf.states(), Z, Zs, "states")
    assert_close(np.stack(inf.covariances()), COV, COVs, "covariances")
    assert np.array_equal(inf.states()[0], np.zeros(k))       # uninferred (ppca_model.rs:98-104)
    assert np.array_equal(inf.covariances()[9], np.eye(k))


@pytest.mark.parametrize("n,d,k,p", SHAPES)
def test_smooth_extrapolate(pk, orc, n, d, k, p):
    X, C0, mu0, s0 = _case(n, d, k, p, empty_rows=(2,))
    w = np.random.default_rng(3).random(n) + 0.1
    ds = pk.Dataset(X, w)
    model = pk.PPCAModel(0.7, C0, mu0)
    sm = model.smooth(ds)
    ex = model.extrapolate(ds)
    want_sm, want_sm_s = both(orc, orc.smooth, X, C0, mu0, 0.7)
    want_ex, want_ex_s = both(orc, orc.extrapolate, X, C0, mu0, 0.7)
    got_sm, got_ex = sm.numpy(), ex.numpy()
    assert np.isfinite(got_sm).all() and np.isfinite(got_ex).all()   # outputs are unmasked datasets
    fin = np.isfinite(X)
    assert np.array_equal(got_ex[fin], X[fin])                        # observed slots bit-identical
    assert_close(got_sm, want_sm, want_sm_s, "smooth")
    assert_close(got_ex, want_ex, want_ex_s, "extrapolate")
    assert np.array_equal(sm.weights(), w) and np.array_equal(ex.weights(), w)  # weights carried (:242,259)
    assert np.array_equal(model.filter_extrapolate(ds).numpy(), got_sm)


@pytest.mark.parametrize("n,d,k,p", SHAPES)
def test_iterate_trajectory(pk, orc, n, d, k, p):
    """Per-iteration llk, C, mu, sigma^2 over 6 EM iterations, each step started from the oracle's model."""
    X, C0, mu0, s0 = _case(n, d, k, p, empty_rows=(4,), empty_dims=(d - 1,) if d > 3 else ())
    w = np.random.default_rng(11).random(n) + 0.5
    ds = pk.Dataset(X, w)
    C, mu, s = C0, mu0, s0
    for it in range(6):
        model = pk.PPCAModel(s, C, mu)
        new, llk = model._iterate(ds, None)
        (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C, mu, s)
        llkw = orc.llk(X, w, C, mu, s)
        assert abs(llk - llkw) <= TOL * abs(llkw), f"llk iteration {it}"
        assert_close(new.transform, Cw, Cs, f"C iteration {it}")
        assert_close(new.mean, muw, mus, f"mu iteration {it}")
        assert_close(new.isotropic_noise ** 2, sw ** 2, ss ** 2, f"sigma^2 iteration {it}")
        if d > 3:  # empty dimension keeps its (zeroed) row and mean (ppca_model.rs:313-321,376)
            assert np.array_equal(new.transform[d - 1], C[d - 1])
        C, mu, s = Cw, muw, sw


def test_iterate_free_running(pk, orc):
    """10 iterations without re-synchronising; llk must not decrease (ppca_model.rs:263-265)."""
    X, C0, mu0, s0 = _case(2000, 60, 8, 0.2, seed=5)
    ds = pk.Dataset(X)
    model = pk.PPCAModel(s0, C0, mu0)
    C, mu, s = C0, mu0, s0
    last = -np.inf
    for it in range(10):
        llk = model.llk(ds)
        assert llk >= last - 1e-9 * abs(llk)
        last = llk
        model = model.iterate(ds)
        C, mu, s = orc.iterate(X, None, C, mu, s)
    assert rel_err(model.transform, C) < 1e-7 and rel_err(model.mean, mu) < 1e-7
    assert abs(model.isotropic_noise - s) < 1e-7 * s
    canon = model.to_canonical()
    assert abs(canon.llk(ds) - model.llk(ds)) < 1e-9 * abs(model.llk(ds))  # ppca_model.rs:395-397


def test_priors(pk, orc):
    X, C0, mu0, s0 = _case(800, 12, 3, 0.25, seed=9)
    ds = pk.Dataset(X)
    rng = np.random.default_rng(2)
    A = rng.standard_normal((12, 12))
    cov = A @ A.T / 12 + np.eye(12)
    m0 = rng.standard_normal(12)
    prior = (pk.Prior().with_mean_prior(m0.reshape(1, -1), cov).with_isotropic_noise_prior(3.0, 2.0)
             .with_transformation_precision(0.5))
    oprior = orc.Prior(mean=m0, mean_covariance=cov, isotropic_noise_alpha=3.0, isotropic_noise_beta=2.0,
                       transformation_precision=0.5)
    C, mu, s = C0, mu0, s0
    for it in range(4):
        new = pk.PPCAModel(s, C, mu).iterate_with_prior(ds, prior)
        (C, mu, s), (Cs, mus, ss) = both(orc, orc.iterate, X, None, C, mu, s, oprior)
        assert_close(new.transform, C, Cs, "C")
        assert_close(new.mean, mu, mus, "mu")
        assert_close(new.isotropic_noise, s, ss, "sigma")


def _mix_case(pk, n, d, ks, seed=0):
    rng = np.random.default_rng(seed)
    X = np.concatenate([make_data(n // len(ks), d, k, 0.2, seed=seed + j, mean_scale=2.0) for j, k in enumerate(ks)])
    rng.shuffle(X, axis=0)
    models = []
    for j, k in enumerate(ks):
        C0, mu0, s0 = init_model(d, k, seed=1000 + j)
        models.append((C0, mu0 + 0.1 * j, 1.0 + 0.1 * j))
    logw = np.log(np.arange(1, len(ks) + 1) / np.sum(np.arange(1, len(ks) + 1)))
    return X, models, logw


@pytest.mark.parametrize("ks", [(2, 2), (4, 4, 4), (3, 5, 2, 8), (34, 12, 48)])
def test_mixture(pk, orc, ks):
    n, d = (900, 24) if max(ks) <= 8 else (1200, 72)   # the last case runs the register-tiled solve inside the mixture pass
    X, models, logw = _mix_case(pk, n, d, ks)
    w = np.random.default_rng(5).random(X.shape[0]) + 0.5
    ds = pk.Dataset(X, w)
    mix = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in models], logw)
    assert rel_err(mix.llks(ds), orc.mix_llks(X, models, logw)) < TOL
    want_llk = orc.mix_llk(X, w, models, logw)
    assert abs(mix.llk(ds) - want_llk) < TOL * abs(want_llk)
    assert np.max(np.abs(mix.infer_cluster(ds) - orc.mix_infer_cluster(X, models, logw))) < 1e-9
    sm, ex = mix.smooth(ds).numpy(), mix.extrapolate(ds).numpy()
    want_sm, want_sm_s = both(orc, orc.mix_smooth, X, models, logw)
    want_ex, want_ex_s = both(orc, orc.mix_smooth, X, models, logw, extrapolate=True)
    assert_close(sm, want_sm, want_sm_s, "mix smooth")
    assert_close(ex, want_ex, want_ex_s, "mix extrapolate")
    assert np.array_equal(mix.smooth(ds).weights(), np.ones(X.shape[0]))     # weights reset (mix.rs:245-265)
    cur_models, cur_logw = models, logw
    for it in range(3):
        m = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in cur_models], cur_logw)
        new, llk = m._iterate(ds, None)
        (want_models, want_logw), (st_models, _) = both(orc, orc.mix_iterate, X, w, cur_models, cur_logw)
        assert abs(llk - orc.mix_llk(X, w, cur_models, cur_logw)) < TOL * abs(llk)
        assert np.max(np.abs(new.log_weights - want_logw)) < 1e-9
        for got, (Cw, muw, sw), (Cs, mus, ss) in zip(new.models, want_models, st_models):
            assert_close(got.transform, Cw, Cs, "mix C")
            assert_close(got.mean, muw, mus, "mix mu")
            assert_close(got.isotropic_noise, sw, ss, "mix sigma")
        cur_models, cur_logw = want_models, want_logw


def test_trainer_and_pickle(pk, capsys):
    import pickle
    X = make_data(500, 10, 3, 0.2, seed=4)
    ds = pk.Dataset(X)
    model = pk.PPCATrainer(ds).train(state_size=3, n_iters=5)
    out = capsys.readouterr().out
    assert out.count("Masked PPCA iteration") == 5 and "aic=" in out
    again = pickle.loads(pickle.dumps(model))
    assert np.array_equal(again.transform, model.transform) and again.isotropic_noise == model.isotropic_noise
    assert pk.PPCAModel.load(model.dump()).llk(ds) == model.llk(ds)
    mix = pk.PPCAMixTrainer(ds).train(n_models=2, state_size=2, n_iters=3, quiet=True)
    assert pickle.loads(pickle.dumps(mix)).llk(ds) == mix.llk(ds)
    assert mix.n_parameters == sum(m.n_parameters for m in mix.models) + 1


# ---- committed golden vectors (tests/golden, made by tests/golden/make_golden.py) -----------------------
@pytest.mark.parametrize("name", ["toy_d3_k2", "ragged_d37_k5", "c2shape_d200_k16"])
def test_golden_single(pk, name):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    X, w = g["X"], g["w"]
    ds = pk.Dataset(X, w)
    model = pk.PPCAModel(float(g["s0"]), g["C0"], g["mu0"])
    assert rel_err(model.llks(ds), g["llks0"]) < TOL
    inf = model.infer(ds)
    assert rel_err(inf.states(), g["Z0"]) < TOL and rel_err(np.stack(inf.covariances()), g["COV0"]) < TOL
    ex = model.extrapolate(ds).numpy()
    fin = np.isfinite(X)
    assert np.array_equal(ex[fin], X[fin]) and rel_err(ex, g["extrapolate0"]) < TOL
    C, mu, s = g["C0"], g["mu0"], float(g["s0"])
    for it in range(g["C_traj"].shape[0]):
        new, llk = pk.PPCAModel(s, C, mu)._iterate(ds, None)
        assert abs(llk - float(g["llk_traj"][it])) < TOL * abs(llk)
        # the golden trajectory carries the reference's own cancellation noise (see assert_close); at these
        # sigma values it stays below ~3e-9
        assert rel_err(new.transform, g["C_traj"][it]) < 5e-9 and rel_err(new.mean, g["mu_traj"][it]) < 5e-9
        assert abs(new.isotropic_noise - float(g["s_traj"][it])) < 5e-9 * new.isotropic_noise
        C, mu, s = g["C_traj"][it], g["mu_traj"][it], float(g["s_traj"][it])


def test_golden_mixture(pk):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mix_d12_k232.npz"))
    ds = pk.Dataset(g["X"], g["w"])
    m = len(g["ks"])
    mix = pk.PPCAMix([pk.PPCAModel(float(g[f"s0_{j}"]), g[f"C0_{j}"], g[f"mu0_{j}"]) for j in range(m)], g["logw0"])
    assert rel_err(mix.llks(ds), g["llks0"]) < TOL
    assert np.max(np.abs(mix.infer_cluster(ds) - g["logpost0"])) < 1e-9
    assert rel_err(mix.extrapolate(ds).numpy(), g["extrapolate0"]) < TOL
    new, llk = mix._iterate(ds, None)
    assert abs(llk - float(g["llk_0"])) < TOL * abs(llk)
    assert np.max(np.abs(new.log_weights - g["logw_0"])) < 1e-9
    for j, got in enumerate(new.models):
        assert rel_err(got.transform, g[f"C_0_{j}"]) < 5e-9


# ---- the sharded path through the C ABI: virtual shards on one GPU --------------------------------------
def test_virtual_shards_equal_single_pass(pk, orc):
    """ppca_b200_em_stats on R row blocks, buffers summed in rank order (what the NCCL all-reduce does),
    then ppca_b200_em_finish == one ppca_b200_iterate on the whole dataset."""
    import torch
    from ppca_rs_b200.distributed import CudaEngine, shard_bounds
    n, d, k = 5000, 90, 12
    X, C0, mu0, s0 = _case(n, d, k, 0.25, seed=21)
    w = np.random.default_rng(2).random(n) + 0.5
    model = pk.PPCAModel(s0, C0, mu0)
    whole = pk.Dataset(X, w)
    want, want_llk = model._iterate(whole, None)
    eng = CudaEngine(pk.get_context())
    for world in (2, 3):
        total = None
        for r in range(world):
            lo, hi = shard_bounds(n, world, r)
            st = eng.new_stats(d, k)
            eng.em_stats(pk.Dataset(X[lo:hi], w[lo:hi]), model, st)
            torch.cuda.synchronize()
            total = st.clone() if total is None else total + st
        got, llk = eng.em_finish(model, None, total)
        assert rel_err(got.transform, want.transform) < 1e-12 and rel_err(got.mean, want.mean) < 1e-12
        assert abs(got.isotropic_noise - want.isotropic_noise) < 1e-12 * want.isotropic_noise
        assert abs(llk - want_llk) < 1e-12 * abs(want_llk)


# ---- out-of-core path: samples streamed from host memory every step ---------------------------------------
@pytest.mark.parametrize("n,d,k,chunk,weighted,pin", [
    (5000, 90, 12, 0, True, True),       # one (tail) block
    (5000, 90, 12, 1024, False, True),   # 4 full blocks + tail, double-buffered H2D
    (4096, 33, 7, 1024, True, False),    # exact multiple of the block, pageable memory
    (700, 200, 16, 256, False, True),
])
def test_iterate_host_streaming_equals_resident(pk, orc, n, d, k, chunk, weighted, pin):
    """ppca_b200_iterate_host (HostDataset) == ppca_b200_dataset_from_host + ppca_b200_iterate, and == oracle."""
    X, C0, mu0, s0 = _case(n, d, k, 0.25, seed=31)
    X[5, :] = np.nan                                   # an empty sample
    w = (np.random.default_rng(3).random(n) + 0.5) if weighted else None
    ctx = pk.Context(0)
    try:
        ctx.set_chunk(chunk)
        model = pk.PPCAModel(s0, C0, mu0)
        host = pk.HostDataset(X, w, pin=pin, ctx=ctx)
        assert len(host) == n and host.output_size() == d
        prior = pk.Prior().with_transformation_precision(0.5)
        for pr in (None, prior):
            want, want_llk = model._iterate(pk.Dataset(X, w, _ctx=ctx), pr)
            for _ in range(2):                         # second call reuses the cached block stores
                got, llk = model._iterate(host, pr)
                assert rel_err(got.transform, want.transform) < 1e-12 and rel_err(got.mean, want.mean) < 1e-12
                assert abs(got.isotropic_noise - want.isotropic_noise) < 1e-12 * want.isotropic_noise
                assert abs(llk - want_llk) < 1e-12 * abs(want_llk)
        (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C0, mu0, s0)
        got, llk = model._iterate(host, None)
        assert_close(got.transform, Cw, Cs, "C")
        assert_close(got.mean, muw, mus, "mu")
        assert_close(got.isotropic_noise ** 2, sw ** 2, ss ** 2, "sigma^2")
        del host
    finally:
        ctx.close()


def test_trainer_on_host_dataset(pk):
    X, C0, mu0, s0 = _case(3000, 40, 4, 0.2, seed=5)
    start = pk.PPCAModel(s0, C0, mu0)
    a = pk.PPCATrainer(pk.Dataset(X)).train(start=start, state_size=4, n_iters=5, quiet=True)
    b = pk.PPCATrainer(pk.HostDataset(X)).train(start=start, state_size=4, n_iters=5, quiet=True)
    assert rel_err(b.transform, a.transform) < 1e-10 and abs(b.isotropic_noise - a.isotropic_noise) < 1e-10


@pytest.mark.parametrize("n,d,k,chunk,pin", [(5000, 90, 12, 1024, True), (3000, 130, 48, 0, True), (2048, 33, 7, 512, False)])
def test_host_streamed_inference_equals_resident(pk, orc, n, d, k, chunk, pin):
    """ppca_b200_reconstruct_host (smooth / extrapolate / llks on a HostDataset) == the resident entry points."""
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=41)
    X[7, :] = np.nan
    w = np.random.default_rng(5).random(n) + 0.5
    ctx = pk.Context(0)
    try:
        ctx.set_chunk(chunk)
        model = pk.PPCAModel(0.3, C0, mu0)
        res, host = pk.Dataset(X, w, _ctx=ctx), pk.HostDataset(X, w, pin=pin, ctx=ctx)
        for _ in range(2):
            ex = model.extrapolate(host)
            assert isinstance(ex, pk.HostDataset) and np.array_equal(ex.weights(), w)
            got, want = ex.numpy(), model.extrapolate(res).numpy()
            fin = np.isfinite(X)
            assert np.array_equal(got[fin], X[fin])                       # observed slots: bit copies
            assert rel_err(got, want) < 1e-13
            assert rel_err(model.smooth(host).numpy(), model.smooth(res).numpy()) < 1e-13
            assert rel_err(model.llks(host), model.llks(res)) < 1e-13
            assert abs(model.llk(host) - model.llk(res)) < 1e-12 * abs(model.llk(res))
        assert rel_err(model.extrapolate(host).numpy(), orc.extrapolate(X, C0, mu0, 0.3)) < 1e-9
    finally:
        ctx.close()


def test_dataframe_adapter_roundtrip_on_device(pk):
    pd = pytest.importorskip("pandas")
    rng = np.random.default_rng(0)
    rows = [{"unit": u, "t": t, "sensor": f"s{j}", "value": float(np.sin(0.1 * t * (j + 1)) + 0.01 * rng.standard_normal())}
            for u in range(6) for t in range(40) for j in range(5) if rng.random() > 0.25]
    df = pd.DataFrame(rows)
    ad = pk.DataFrameAdapter.from_pandas(df, keys=["unit", "t"], dimensions=["sensor"], metric="value")
    assert isinstance(ad.dataset, pk.Dataset) and ad.dataset.output_size() == 5
    model = pk.PPCATrainer(ad.dataset).train(state_size=2, n_iters=8, quiet=True)
    long = ad.convert_datasets({"value_hat": model.extrapolate(ad.dataset), "value_in": ad.dataset})
    assert len(long) == len(ad.sample_idx) * 5 and not long["value_hat"].isna().any()
    merged = long.merge(df, on=["unit", "t", "sensor"], how="left")
    obs = ~merged["value"].isna()
    assert np.array_equal(merged.loc[obs, "value_hat"].to_numpy(), merged.loc[obs, "value"].to_numpy())
    assert merged.loc[~obs, "value_in"].isna().all()
    host = pk.DataFrameAdapter.from_pandas(df, keys=["unit", "t"], dimensions=["sensor"], metric="value",
                                           dataset_factory=pk.HostDataset)
    assert np.allclose(host.dataset.numpy(), ad.dataset.numpy(), equal_nan=True)


def test_model_sample_on_device_has_the_model_distribution(pk):
    """PPCAModel.sample (ppca_model.rs:164-191): x ~ N(mu, C C^T + sigma^2 I), entries masked i.i.d. with prob p.
    The reference is unseeded, so parity is distributional; a given seed reproduces."""
    rng = np.random.default_rng(3)
    d, k, n, p = 12, 3, 400_000, 0.3
    C0, mu0, s0 = rng.standard_normal((d, k)), rng.standard_normal(d), 0.5
    model = pk.PPCAModel(s0, C0, mu0)
    a = model.sample(n, p, seed=11).numpy()
    assert a.shape == (n, d)
    fin = np.isfinite(a)
    assert abs(fin.mean() - (1 - p)) < 5e-3
    cov_true = C0 @ C0.T + s0 ** 2 * np.eye(d)
    mean = np.nanmean(a, axis=0)
    assert np.max(np.abs(mean - mu0)) < 6 * np.sqrt(np.max(np.diag(cov_true)) / (n * (1 - p)))
    xc = np.where(fin, a - mu0, 0.0)
    pair = fin.astype(np.float64).T @ fin.astype(np.float64)
    cov = (xc.T @ xc) / pair                              # pairwise-complete covariance
    assert np.max(np.abs(cov - cov_true)) < 0.05 * np.max(np.abs(cov_true))
    assert np.array_equal(model.sample(1000, p, seed=11).numpy(), a[:1000], equal_nan=True)   # counter-based RNG
    assert not np.array_equal(model.sample(1000, p, seed=12).numpy(), a[:1000], equal_nan=True)
    assert len(model.sample(0, p)) == 0
    fresh = model.sample(50_000, 0.2, seed=5)
    fitted = pk.PPCATrainer(fresh).train(state_size=k, n_iters=60, quiet=True)
    assert abs(fitted.isotropic_noise - s0) < 0.05        # (the mean converges slowly under EM from a random start)
    assert fitted.llk(fresh) > model.llk(fresh) - 0.01 * abs(model.llk(fresh))   # the fit explains the draw like the truth


@pytest.mark.parametrize("n,d,k", [(3000, 37, 5), (2500, 200, 16), (700, 130, 48), (300, 20, 1)])
def test_covariance_diagonals_on_device(pk, n, d, k):
    """InferredMasked.smoothed_/extrapolated_covariances_diagonal (ppca_model.rs:485-508, 542-577) on the device
    (ppca_b200_covariance_diagonal) against the defining formula diag(sigma^2 I + C Sigma_n C^T)."""
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=51)
    X[3, :] = np.nan
    model = pk.PPCAModel(0.4, C0, mu0)
    ds = pk.Dataset(X)
    inf = model.infer(ds)
    covs = np.stack(inf.covariances())
    want = np.einsum("ia,nab,ib->ni", C0, covs, C0) + 0.4 ** 2
    got = inf.smoothed_covariances_diagonal(model)
    assert isinstance(got, pk.Dataset) and got.empty_dimensions() == []
    assert rel_err(got.numpy(), want) < 1e-12
    ex = inf.extrapolated_covariances_diagonal(model, ds).numpy()
    fin = np.isfinite(X)
    assert np.all(ex[fin] == 0.0) and rel_err(ex[~fin], want[~fin]) < 1e-12
    full = inf.smoothed_covariances(model)                      # the d x d matrices (host) share the diagonal
    assert rel_err(np.diag(full[0]), want[0]) < 1e-12
    # mixture: law of total variance over the components (mix.rs:445-458)
    if k <= 16:
        other = pk.PPCAModel(0.7, C0[:, ::-1].copy(), mu0 + 0.3)
        mix = pk.PPCAMix([model, other], np.log([0.4, 0.6]))
        im = mix.infer(ds)
        post = im.posteriors()
        means = [im._inf[j]._states @ m.transform.T + m.mean for j, m in enumerate([model, other])]
        mean = sum(post[:, j:j + 1] * means[j] for j in range(2))
        diags = [np.einsum("ia,nab,ib->ni", m.transform, np.stack(im._inf[j].covariances()), m.transform)
                 + m.isotropic_noise ** 2 for j, m in enumerate([model, other])]
        want_mix = sum(post[:, j:j + 1] * (diags[j] + (means[j] - mean) ** 2) for j in range(2))
        assert rel_err(im.smoothed_covariances_diagonal(mix).numpy(), want_mix) < 1e-11


# ---- the reference's own smoke tests (ppca/src/lib.rs:47-105), same flow through this API, with assertions ----------
def test_reference_toy_model_flow(pk):
    """lib.rs:47-62 test_toy_model: sample 1000 draws (20 % masked) from the 3 x 2 toy model, init(2), 1600 iterations."""
    real = pk.PPCAModel(0.1, np.array([[1.0, 1.0], [1.0, 0.0], [0.0, 1.0]]), np.array([[0.0], [1.0], [0.0]]))
    sample = real.sample(1_000, 0.2, seed=7)
    model = pk.PPCAModel.init(2, sample, seed=1)
    aic = []
    for _ in range(1600):
        aic.append(2.0 * (model.n_parameters - model.llk(sample)) / len(sample))
        model = model.iterate(sample)
    assert all(b <= a + 1e-9 * abs(a) for a, b in zip(aic, aic[1:]))          # EM never increases the AIC
    assert abs(model.isotropic_noise - 0.1) < 0.02
    assert model.llk(sample) >= real.llk(sample) - 1e-6 * abs(real.llk(sample))   # the MLE beats the truth on its sample


def test_reference_big_toy_model_flow(pk):
    """lib.rs:82-101 test_big_toy_model: 200 x 16 Bernoulli(0.1) transform, sigma 0.1, 100 000 draws (20 % masked),
    init(16), 24 iterations, to_canonical."""
    rng = np.random.default_rng(0)
    real = pk.PPCAModel(0.1, (rng.random((200, 16)) < 0.1).astype(np.float64), np.zeros(200))
    sample = real.sample(100_000, 0.2, seed=3)
    model = pk.PPCAModel.init(16, sample, seed=2)
    aic = []
    for _ in range(24):
        aic.append(2.0 * (model.n_parameters - model.llk(sample)) / len(sample))
        model = model.iterate(sample)
    assert all(b <= a + 1e-9 * abs(a) for a, b in zip(aic, aic[1:]))
    canon = model.to_canonical()
    assert abs(canon.llk(sample) - model.llk(sample)) < 1e-9 * abs(model.llk(sample))
    sv = canon.singular_values
    assert np.all(np.diff(sv) <= 1e-12) and canon.isotropic_noise < 0.5


@pytest.mark.parametrize("script", ["toy_model.py", "big_toy_model.py", "ppca_mixture.py",
                                    "priors_pickling_empty_dimensions.py", "device_resident_and_out_of_core.py"])
def test_examples_run_unmodified_ppca_rs_calls(script, capsys):
    """examples/ use the reference's import lines (`from ppca_rs import ...`, `from ppca_rs.ppca_rs import ...`) through
    ppca_rs_b200.compat.install_as_ppca_rs()."""
    import os
    import runpy
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "examples"))
    try:
        runpy.run_path(os.path.join(root, "examples", script), run_name="__main__")
    finally:
        sys.path.remove(os.path.join(root, "examples"))
        for name in ("ppca_rs", "ppca_rs.ppca_rs", "_alias"):
            sys.modules.pop(name, None)
    assert capsys.readouterr().out.strip() != ""


# ---- full-size properties (BASELINE configs[1]: N=1M, d=200, k=16, 20% missing) ---------------------------
def test_full_size_properties(pk):
    n, d, k = 1_000_000, 200, 16
    ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.2, seed=20240531)
    assert len(ds) == n and ds.output_size() == d and ds.empty_dimensions() == []
    rng = np.random.default_rng(1)
    model = pk.PPCAModel(1.0, rng.standard_normal((d, k)), np.zeros(d))
    llks = []
    for _ in range(4):
        model, llk = model._iterate(ds, None)
        llks.append(llk)
    assert all(b >= a for a, b in zip(llks, llks[1:]))                  # EM never decreases the llk (:263-265)
    assert abs(model.llk(ds) - float(np.sum(model.llks(ds)))) < 1e-9 * abs(llks[-1])   # llk == sum of llks
    canon = model.to_canonical()
    assert abs(canon.llk(ds) - model.llk(ds)) < 1e-9 * abs(llks[-1])   # :395-397
    head = ds._slice(0, 4096)
    x = head.numpy()
    ex = model.extrapolate(head).numpy()
    fin = np.isfinite(x)
    assert 0.78 < fin.mean() < 0.82                                     # Bernoulli(0.8) mask
    assert np.array_equal(ex[fin], x[fin]) and np.isfinite(ex).all()    # observed slots untouched (:246-247)
    sm = model.smooth(head).numpy()
    assert np.array_equal(ex[~fin], sm[~fin])                           # extrapolate = choose(x, smoothed)
    # idempotence: smoothing a fully observed reconstruction of rank k with tiny noise moves it very little
    assert np.sqrt(np.mean((sm[fin] - x[fin]) ** 2)) < 5 * model.isotropic_noise


# ---- the exact int8-sliced evaluation of the two masked-Gram contractions (csrc/ibitgemm.cu) -----------------
@pytest.fixture
def int8_ctx(pk):
    ctx = pk.get_context()
    yield ctx
    ctx.set_gemm("tc", 6)  # the engine default


@pytest.mark.parametrize("mode", ["dmma", "int8", "tc"])
@pytest.mark.parametrize("slices", [6, 7])
@pytest.mark.parametrize("n,d,k,p", [(100, 3, 2, 0.2), (777, 37, 5, 0.3), (3000, 200, 16, 0.2), (1200, 150, 32, 0.25),
                                     (500, 260, 64, 0.3), (300, 20, 1, 0.1), (20000, 70, 10, 0.25)])
def test_int8_sliced_iterate(pk, orc, int8_ctx, mode, slices, n, d, k, p):
    int8_ctx.set_gemm(mode, slices)
    X, C0, mu0, s0 = _case(n, d, k, p, empty_rows=(4,), empty_dims=(d - 1,) if d > 3 else ())
    w = np.random.default_rng(11).random(n) + 0.5
    ds = pk.Dataset(X, w)
    model = pk.PPCAModel(0.4, C0, mu0)
    assert rel_err(model.llks(ds), orc.llks(X, C0, mu0, 0.4)) < TOL
    (Z, COV), (Zs, COVs) = both(orc, orc.infer, X, C0, mu0, 0.4)
    inf = model.infer(ds)
    assert_close(inf.states(), Z, Zs, "states")
    assert_close(np.stack(inf.covariances()), COV, COVs, "covariances")
    C, mu, s = C0, mu0, s0
    for it in range(5):
        new, llk = pk.PPCAModel(s, C, mu)._iterate(ds, None)
        (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C, mu, s)
        llkw = orc.llk(X, w, C, mu, s)
        assert abs(llk - llkw) <= TOL * abs(llkw), f"llk iteration {it}"
        assert_close(new.transform, Cw, Cs, f"C iteration {it}")
        assert_close(new.mean, muw, mus, f"mu iteration {it}")
        assert_close(new.isotropic_noise ** 2, sw ** 2, ss ** 2, f"sigma^2 iteration {it}")
        C, mu, s = Cw, muw, sw


@pytest.mark.parametrize("n,d,k,p", [(777, 37, 5, 0.3), (3000, 200, 16, 0.2), (1200, 150, 32, 0.25), (500, 260, 64, 0.3)])
def test_fp32_class_fast_path(pk, orc, int8_ctx, n, d, k, p):
    """The opt-in fast path (4 digit planes on tcgen05): BASELINE.json north_star allows 1e-4 relative for the FP32
    path on per-iteration llk, C, mu, sigma^2; masks and observed slots of extrapolate stay bit-identical."""
    X, C0, mu0, s0 = _case(n, d, k, p, seed=61)
    int8_ctx.set_gemm("tc", 4)
    ds = pk.Dataset(X)
    C, mu, s = C0, mu0, s0
    for _ in range(3):
        new, llk = pk.PPCAModel(s, C, mu)._iterate(ds, None)
        with orc.stable():
            Cw, muw, sw = orc.iterate(X, None, C, mu, s)
            llkw = orc.llk(X, None, C, mu, s)
        assert abs(llk - llkw) < 1e-4 * abs(llkw)
        assert rel_err(new.transform, Cw) < 1e-4 and rel_err(new.mean, muw) < 1e-4
        assert abs(new.isotropic_noise ** 2 - sw ** 2) < 1e-4 * sw ** 2
        C, mu, s = Cw, muw, sw
    ex = pk.PPCAModel(s, C, mu).extrapolate(ds).numpy()
    fin = np.isfinite(X)
    assert np.array_equal(ex[fin], X[fin]) and np.isfinite(ex).all()
    with pytest.raises(Exception):
        int8_ctx.set_gemm("int8", 4)          # the fast path exists on the tcgen05 mode only


def test_int8_sliced_matches_dmma_closely(pk, int8_ctx):
    """Same inputs through both arithmetic paths: the int8-sliced statistics agree with DMMA to ~1e-13."""
    n, d, k = 20000, 200, 16
    ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.2, seed=5)
    rng = np.random.default_rng(2)
    model = pk.PPCAModel(1.0, rng.standard_normal((d, k)), np.zeros(d))
    int8_ctx.set_gemm("dmma")
    a, llk_a = model._iterate(ds, None)
    for mode, slices, tol in (("int8", 8, 5e-13), ("int8", 7, 5e-13), ("int8", 6, 5e-13),
                              ("tc", 8, 5e-13), ("tc", 7, 5e-13), ("tc", 6, 5e-13)):
        int8_ctx.set_gemm(mode, slices)
        b, llk_b = model._iterate(ds, None)
        assert rel_err(b.transform, a.transform) < tol and rel_err(b.mean, a.mean) < tol, (mode, slices)
        assert abs(b.isotropic_noise - a.isotropic_noise) < tol * a.isotropic_noise
        assert abs(llk_b - llk_a) < tol * abs(llk_a)


def test_dmma_mixture_and_inference(pk, orc, int8_ctx):
    """Mixture EM and extrapolate with the FP64 DMMA contraction (the other tests run the tcgen05 default)."""
    int8_ctx.set_gemm("dmma")
    X, models, logw = _mix_case(pk, 900, 24, (3, 5, 2, 8))
    w = np.random.default_rng(5).random(X.shape[0]) + 0.5
    ds = pk.Dataset(X, w)
    mix = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in models], logw)
    assert rel_err(mix.llks(ds), orc.mix_llks(X, models, logw)) < TOL
    want_ex, want_ex_s = both(orc, orc.mix_smooth, X, models, logw, extrapolate=True)
    assert_close(mix.extrapolate(ds).numpy(), want_ex, want_ex_s, "mix extrapolate")
    new, llk = mix._iterate(ds, None)
    (want_models, want_logw), (st_models, _) = both(orc, orc.mix_iterate, X, w, models, logw)
    assert np.max(np.abs(new.log_weights - want_logw)) < 1e-9
    for got, (Cw, muw, sw), (Cs, mus, ss) in zip(new.models, want_models, st_models):
        assert_close(got.transform, Cw, Cs, "mix C")
        assert_close(got.isotropic_noise, sw, ss, "mix sigma")


# ---- the c3 / c5 contraction kernel (tbitgemm_atm2_kernel: two output tiles per expanded mask stage) ---------------
# It is the default whenever a tile is > 4 K steps and the paired tile count balances the persistent grid (d > 512 in
# the E-step; the M-step when n / 128 > 4 K steps per slab and ceil(d/128) * ceil(qtiles/2) * splitk fills 148 CTAs).
# The per-variant launch counters prove which kernel ran.
@pytest.mark.parametrize("n,d,k,p,e_atm2,m_atm2", [
    (4096, 2048, 64, 0.3, True, True),     # BASELINE configs[2] shape: both contractions on the two-tile kernel
    (8192, 1024, 48, 0.3, True, False),    # BASELINE configs[4] shape: E-step two-tile, M-step one-tile feed kernel
    (5500, 640, 40, 0.25, True, None),     # ragged: kk = 820 (last pair holds one q tile), partial last row tile
])
def test_two_tile_contraction_kernel_parity(pk, orc, n, d, k, p, e_atm2, m_atm2):
    X, C0, mu0, s0 = _case(n, d, k, p, seed=21, empty_rows=(3,), empty_dims=(d - 2,))
    w = np.random.default_rng(5).random(n) + 0.5
    ds = pk.Dataset(X, w)
    ctx = pk.get_context()
    c0 = ctx.variant_counts()
    model = pk.PPCAModel(0.6, C0, mu0)
    got = model.llks(ds)
    c1 = ctx.variant_counts()
    assert (c1["tc_atm2"] - c0["tc_atm2"] >= 1) == e_atm2, (c0, c1)
    want = orc.llks(X, C0, mu0, 0.6)
    assert rel_err(got, want) < TOL
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < TOL       # element-wise, not only max-norm
    (Z, COV), (Zs, COVs) = both(orc, orc.infer, X[:512], C0, mu0, 0.6)
    inf = model.infer(ds._slice(0, 512))
    assert_close(inf.states(), Z, Zs, "states")
    assert_close(np.stack(inf.covariances()), COV, COVs, "covariances")
    C, mu, s = C0, mu0, s0
    for it in range(2):
        c2 = ctx.variant_counts()
        new, llk = pk.PPCAModel(s, C, mu)._iterate(ds, None)
        c3 = ctx.variant_counts()
        n_atm2 = c3["tc_atm2"] - c2["tc_atm2"]
        if m_atm2 is not None:
            assert n_atm2 == (1 if e_atm2 else 0) + (1 if m_atm2 else 0), (c2, c3)
        (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C, mu, s)
        llkw = orc.llk(X, w, C, mu, s)
        assert abs(llk - llkw) <= TOL * abs(llkw), f"llk iteration {it}"
        assert_close(new.transform, Cw, Cs, f"C iteration {it}")
        assert_close(new.mean, muw, mus, f"mu iteration {it}")
        assert_close(new.isotropic_noise ** 2, sw ** 2, ss ** 2, f"sigma^2 iteration {it}")
        assert np.array_equal(new.transform[d - 2], C[d - 2])                     # empty dimension keeps its row
        C, mu, s = Cw, muw, sw
    ex = pk.PPCAModel(s, C, mu).extrapolate(ds._slice(0, 1024)).numpy()
    fin = np.isfinite(X[:1024])
    assert np.array_equal(ex[fin], X[:1024][fin])
    want_ex, want_ex_s = both(orc, orc.extrapolate, X[:1024], C, mu, s)
    assert_close(ex, want_ex, want_ex_s, "extrapolate")


def test_posterior_samplers_carry_their_models(pk):
    """InferredMasked.posterior_sampler().sample() takes no arguments (src/python_bindings.rs:335-361, :875-901)."""
    X, C0, mu0, s0 = _case(400, 12, 3, 0.2, seed=2)
    ds = pk.Dataset(X)
    model = pk.PPCAModel(0.5, C0, mu0)
    smp = model.infer(ds).posterior_sampler().sample()
    assert len(smp) == 400 and np.isfinite(smp.numpy()).all()
    mix = pk.PPCAMix([model, pk.PPCAModel(0.7, C0 * 0.5, mu0 + 0.2)], np.log([0.4, 0.6]))
    smp = mix.infer(ds).posterior_sampler().sample()
    assert len(smp) == 400 and np.isfinite(smp.numpy()).all()


def test_slice_recomputes_its_minimum_weight(pk):
    """A chunk with only positive weights from a parent holding a zero weight elsewhere is accepted by mixture EM."""
    X, C0, mu0, s0 = _case(600, 10, 2, 0.2, seed=4)
    w = np.ones(600)
    w[5] = 0.0
    ds = pk.Dataset(X, w)
    mix = pk.PPCAMix([pk.PPCAModel(1.0, C0, mu0), pk.PPCAModel(1.0, -C0, mu0 + 0.1)], np.log([0.5, 0.5]))
    with pytest.raises(Exception):
        mix.iterate(ds)                         # zero weight: mix.rs:304-309
    tail = ds._slice(300, 300)
    assert isinstance(mix.iterate(tail), pk.PPCAMix)
    with pytest.raises(Exception):
        mix.iterate(ds._slice(0, 300))


def test_two_contexts_on_two_devices_in_one_process(pk):
    """cudaFuncSetAttribute is per device: a second context on another GPU must opt its kernels in again."""
    from ppca_rs_b200 import _native as nat
    if nat.device_count() < 2:
        pytest.skip("needs two GPUs")
    X, C0, mu0, s0 = _case(3000, 200, 16, 0.2)
    outs = []
    for dev in (0, 1):
        ctx = nat.Context(dev)
        ds = pk.Dataset(X, _ctx=ctx)
        new, llk = pk.PPCAModel(s0, C0, mu0)._iterate(ds, None)
        outs.append((new.transform, llk))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


# ---- precision guard of the int8-sliced contractions: adversarial dynamic range -----------------------------------
# The default arithmetic keeps every term to 46 bits below its COLUMN maximum.  These cases put the column maximum far
# above what some samples / dimensions see; the guard must notice and the ladder (8 planes, then FP64 DMMA) must bring
# the result back to the accuracy of an FP64 evaluation.  `floor` = what two FP64 evaluations of the same quantity
# (the oracle and the independent dense restatement tests/dense_ref.py) disagree by on these ill-scaled inputs.
def _elementwise(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))


@pytest.fixture
def guard_ctx(pk):
    ctx = pk.get_context()
    ctx.set_gemm("tc", 6)
    ctx.set_guard(True, 40)
    yield ctx
    ctx.set_gemm("tc", 6)
    ctx.set_guard(True, 40)


@pytest.mark.parametrize("scale", [1e2, 1e3, 1e4])
def test_precision_guard_one_feature_in_other_units(pk, orc, guard_ctx, scale):
    import dense_ref
    n, d, k = 1500, 40, 5
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=3)
    X[:, 7] *= scale
    C0 = C0.copy()
    C0[7] *= scale
    ds = pk.Dataset(X)
    model = pk.PPCAModel(0.5, C0, mu0)
    want = orc.llks(X, C0, mu0, 0.5)
    dense = np.array([dense_ref.llk_one(X[i], C0, mu0, 0.5) for i in range(n)])
    floor = _elementwise(dense, want)
    guard_ctx.set_gemm("dmma")
    fp64 = model.llks(ds)
    guard_ctx.set_gemm("tc", 6)
    before = guard_ctx.variant_counts()["precision_retry"]
    got = model.llks(ds)
    retried = guard_ctx.variant_counts()["precision_retry"] - before
    guard_ctx.set_guard(False)
    raw = model.llks(ds)                                  # what the unguarded 46-bit path would have returned
    guard_ctx.set_guard(True)
    print(f"scale {scale:g}: guarded {_elementwise(got, want):.2e} unguarded {_elementwise(raw, want):.2e} "
          f"fp64 {_elementwise(fp64, want):.2e} floor {floor:.2e} retries {retried}")
    assert retried >= 1
    assert _elementwise(got, want) < TOL + 4.0 * floor
    assert _elementwise(got, fp64) < TOL + 4.0 * floor
    # one EM iteration from the ill-scaled model
    guard_ctx.set_gemm("tc", 6)
    new, llk = pk.PPCAModel(1.0, C0, mu0)._iterate(ds, None)
    with orc.stable():
        Cs, mus, ss = orc.iterate(X, None, C0, mu0, 1.0)
    Cd, mud, sd = dense_ref.iterate(X[:400], np.ones(400), C0, mu0, 1.0)
    with orc.stable():
        Cs4, mus4, ss4 = orc.iterate(X[:400], None, C0, mu0, 1.0)
    fl = max(rel_err(Cd, Cs4), rel_err(mud, mus4), abs(sd - ss4) / ss4)
    assert rel_err(new.transform, Cs) < TOL + 4.0 * fl
    assert rel_err(new.mean, mus) < TOL + 4.0 * fl
    assert abs(new.isotropic_noise - ss) < (TOL + 4.0 * fl) * ss


def test_precision_guard_stays_quiet_on_well_scaled_data(pk, orc, guard_ctx):
    X, C0, mu0, s0 = _case(3000, 200, 16, 0.2)
    ds = pk.Dataset(X)
    before = guard_ctx.variant_counts()
    model = pk.PPCAModel(s0, C0, mu0)
    for _ in range(3):
        model, _ = model._iterate(ds, None)
    model.llks(ds)
    model.extrapolate(ds)
    after = guard_ctx.variant_counts()
    assert after["precision_retry"] == before["precision_retry"]
    assert after["dmma"] == before["dmma"] and after["tc_smem_a"] == before["tc_smem_a"]


def test_precision_guard_weights_with_a_wide_dynamic_range(pk, orc, guard_ctx):
    """A dimension observed only by samples of weight 1e-12: its second-moment matrix A_i is a sum of terms 1e-12
    below the column scale of W (M-step guard)."""
    n, d, k = 2000, 30, 4
    X, C0, mu0, s0 = _case(n, d, k, 0.2, seed=8)
    w = np.ones(n)
    w[:200] = 1e-12
    X[200:, 11] = np.nan                     # dimension 11 is seen by the 200 light samples only
    X[:200, 11] = X[:200, 11] + 1.0
    ds = pk.Dataset(X, w)
    before = guard_ctx.variant_counts()["precision_retry"]
    new, llk = pk.PPCAModel(s0, C0, mu0)._iterate(ds, None)
    assert guard_ctx.variant_counts()["precision_retry"] > before
    (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C0, mu0, s0)
    assert_close(new.transform, Cw, Cs, "C")
    assert np.max(np.abs(new.transform[11] - Cs[11])) < TOL * np.max(np.abs(Cs[11]))   # the light row itself
    assert_close(new.mean, muw, mus, "mu")
    assert_close(new.isotropic_noise ** 2, sw ** 2, ss ** 2, "sigma^2")
    guard_ctx.set_guard(False)
    raw, _ = pk.PPCAModel(s0, C0, mu0)._iterate(ds, None)
    print("unguarded row error", np.max(np.abs(raw.transform[11] - Cs[11])) / np.max(np.abs(Cs[11])))


def test_precision_guard_mixture_responsibilities_spanning_1e30(pk, orc, guard_ctx):
    """Two far-apart clusters: the responsibilities of a component for the other cluster's samples are ~1e-30 and one
    dimension is observed by that other cluster only (mix.rs:304-326)."""
    rng = np.random.default_rng(12)
    n, d = 1600, 16
    A = make_data(n // 2, d, 3, 0.2, seed=1, mean_scale=0.2)
    B = make_data(n // 2, d, 3, 0.2, seed=2, mean_scale=0.2) + 1.5
    A[:, 5] = np.nan                           # dimension 5: cluster B only
    X = np.concatenate([A, B])
    X = X[rng.permutation(n)]
    models = []
    for j, shift in enumerate((0.0, 1.5)):
        C0, mu0, _ = init_model(d, 3, seed=50 + j)
        models.append((0.3 * C0, mu0 + shift, 0.6))
    logw = np.log([0.5, 0.5])
    ds = pk.Dataset(X)
    mix = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in models], logw)
    post = mix.infer_cluster(ds)
    assert np.nanmin(post[np.isfinite(post)]) < np.log(1e-30)          # the case really spans > 30 orders of magnitude
    new, llk = mix._iterate(ds, None)
    (want_models, want_logw), (st_models, _) = both(orc, orc.mix_iterate, X, None, models, logw)
    assert abs(llk - orc.mix_llk(X, None, models, logw)) < TOL * abs(llk)
    assert np.max(np.abs(new.log_weights - want_logw)) < 1e-9
    for got, (Cw, muw, sw), (Cs, mus, ss) in zip(new.models, want_models, st_models):
        assert_close(got.transform, Cw, Cs, "mix C")
        assert np.max(np.abs(got.transform[5] - Cs[5])) < TOL * max(np.max(np.abs(Cs[5])), 1e-300)
        assert_close(got.mean, muw, mus, "mix mu")
        assert_close(got.isotropic_noise, sw, ss, "mix sigma")


# ---- single-pass mixture EM: several chunks, the running per-component maximum is raised by later chunks -----------
@pytest.mark.parametrize("chunk", [512, 1024, 0])
def test_mixture_single_pass_rescales_across_chunks(pk, orc, chunk):
    X, models, logw = _mix_case(pk, 2400, 24, (3, 5, 2, 8), seed=4)
    n = X.shape[0]
    w = np.random.default_rng(5).random(n) + 0.5
    w[: n // 2] *= 1e-3                                   # light samples first: every later chunk raises ln w + lp
    ds = pk.Dataset(X, w)
    ctx = pk.get_context()
    ctx.set_chunk(chunk)
    try:
        mix = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in models], logw)
        new, llk = mix._iterate(ds, None)
    finally:
        ctx.set_chunk(0)
    (want_models, want_logw), (st_models, _) = both(orc, orc.mix_iterate, X, w, models, logw)
    assert abs(llk - orc.mix_llk(X, w, models, logw)) < TOL * abs(llk)
    assert np.max(np.abs(new.log_weights - want_logw)) < 1e-9
    for got, (Cw, muw, sw), (Cs, mus, ss) in zip(new.models, want_models, st_models):
        assert_close(got.transform, Cw, Cs, "mix C")
        assert_close(got.mean, muw, mus, "mix mu")
        assert_close(got.isotropic_noise, sw, ss, "mix sigma")


# ---- the collective behind the C ABI (ppca_b200_comm_* and the *_sharded entry points) ----------------------------
def test_sharded_entry_points_world_of_one(pk, orc):
    """NCCL bound at run time, a communicator of one rank: the sharded calls must reproduce the plain ones bit for bit."""
    from ppca_rs_b200 import _native as nat
    ctx = nat.Context(0)
    ctx.comm_init(nat.Context.comm_unique_id(), 0, 1)
    X, C0, mu0, s0 = _case(3000, 60, 8, 0.2, seed=6)
    w = np.random.default_rng(1).random(3000) + 0.5
    ds = pk.Dataset(X, w, _ctx=ctx)
    model = pk.PPCAModel(s0, C0, mu0)
    a, llk_a = model._iterate(ds, None)
    b, llk_b = model._iterate(ds, None, sharded=True)
    assert np.array_equal(a.transform, b.transform) and np.array_equal(a.mean, b.mean)
    assert a.isotropic_noise == b.isotropic_noise and llk_a == llk_b
    host = pk.HostDataset(X, w, pin=False, ctx=ctx)
    c, llk_c = model._iterate(host, None, sharded=True)
    assert rel_err(c.transform, a.transform) < 1e-12 and abs(llk_c - llk_a) < 1e-12 * abs(llk_a)
    Xm, models, logw = _mix_case(pk, 900, 24, (3, 5, 2))
    dsm = pk.Dataset(Xm, _ctx=ctx)
    mix = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in models], logw)
    m1, l1 = mix._iterate(dsm, None)
    m2, l2 = mix._iterate(dsm, None, sharded=True)
    assert l1 == l2 and np.array_equal(m1.log_weights, m2.log_weights)
    for x, y in zip(m1.models, m2.models):
        assert np.array_equal(x.transform, y.transform)
    ctx.comm_destroy()


def _two_gpu_rank(rank, uid, X, w, C0, mu0, s0, Xm, models, logw, out):
    import ppca_rs_b200 as pk
    from ppca_rs_b200 import _native as nat
    from ppca_rs_b200.distributed import shard_bounds
    ctx = nat.Context(rank)
    ctx.comm_init(uid, rank, 2)
    lo, hi = shard_bounds(X.shape[0], 2, rank)
    ds = pk.Dataset(X[lo:hi], w[lo:hi], _ctx=ctx)
    model = pk.PPCAModel(s0, C0, mu0)
    llks = []
    for _ in range(3):
        model, llk = model._iterate(ds, None, sharded=True)
        llks.append(llk)
    lo, hi = shard_bounds(Xm.shape[0], 2, rank)
    mix = pk.PPCAMix([pk.PPCAModel(s, C, mu) for C, mu, s in models], logw)
    mix, mllk = mix._iterate(pk.Dataset(Xm[lo:hi], _ctx=ctx), None, sharded=True)
    out[rank] = (model, llks, mix, mllk)
    ctx.synchronize()


def test_sharded_entry_points_two_gpus(pk, orc):
    """Two ranks (threads, one GPU each) through ppca_b200_iterate_sharded / ppca_b200_mix_iterate_sharded against the
    oracle on the whole dataset."""
    import threading
    from ppca_rs_b200 import _native as nat
    if nat.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, d, k = 6000, 90, 12
    X, C0, mu0, s0 = _case(n, d, k, 0.25, seed=21)
    w = np.random.default_rng(2).random(n) + 0.5
    Xm, models, logw = _mix_case(pk, 1800, 24, (3, 5, 2, 8), seed=3)
    uid = nat.Context.comm_unique_id()
    out = {}
    th = [threading.Thread(target=_two_gpu_rank, args=(r, uid, X, w, C0, mu0, s0, Xm, models, logw, out)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert set(out) == {0, 1}
    (m0, l0, x0, ml0), (m1, l1, x1, ml1) = out[0], out[1]
    assert np.array_equal(m0.transform, m1.transform) and l0 == l1 and ml0 == ml1     # replicated finish
    C, mu, s = C0, mu0, s0
    with orc.stable():
        for it in range(3):
            assert abs(l0[it] - orc.llk(X, w, C, mu, s)) < TOL * abs(l0[it])
            C, mu, s = orc.iterate(X, w, C, mu, s)
        want_models, want_logw = orc.mix_iterate(Xm, None, models, logw)
    assert rel_err(m0.transform, C) < 1e-8 and rel_err(m0.mean, mu) < 1e-8 and abs(m0.isotropic_noise - s) < 1e-8 * s
    assert abs(ml0 - orc.mix_llk(Xm, None, models, logw)) < TOL * abs(ml0)
    assert np.max(np.abs(x0.log_weights - want_logw)) < 1e-9
    for got, (Cw, muw, sw) in zip(x0.models, want_models):
        assert rel_err(got.transform, Cw) < TOL and rel_err(got.mean, muw) < TOL and abs(got.isotropic_noise - sw) < TOL * sw


def test_reconstruct_one_pass_and_in_place_reuse(pk, orc):
    """extrapolate / smooth + llks from one E-step (ppca_b200_reconstruct), output dataset overwritten in place."""
    X, C0, mu0, s0 = _case(3000, 70, 10, 0.25, seed=13, empty_rows=(7,))
    w = np.random.default_rng(4).random(3000) + 0.5
    ds = pk.Dataset(X, w)
    model = pk.PPCAModel(0.7, C0, mu0)
    ex, ll = model.reconstruct(ds, True, with_llks=True)
    assert np.array_equal(ex.numpy(), model.extrapolate(ds).numpy())
    assert np.array_equal(ll, model.llks(ds)) and np.array_equal(ex.weights(), w)
    assert rel_err(ll, orc.llks(X, C0, mu0, 0.7)) < TOL
    # reuse: another model, another input of the same shape, same output handle
    X2 = make_data(3000, 70, 10, 0.4, seed=99)
    ds2 = pk.Dataset(X2)
    model2 = pk.PPCAModel(0.3, 0.5 * C0, mu0 + 0.1)
    again, ll2 = model2.reconstruct(ds2, False, out=ex, with_llks=True)
    assert again is ex
    assert np.array_equal(ex.numpy(), model2.smooth(ds2).numpy()) and np.array_equal(ll2, model2.llks(ds2))
    assert np.array_equal(ex.weights(), np.ones(3000))
    with pytest.raises(Exception):
        model.reconstruct(ds, True, out=ds2)            # not an output dataset
    with pytest.raises(Exception):
        model.reconstruct(ds._slice(0, 100), True, out=ex)   # another shape


def test_mean_prior_device_cholesky_matches_the_reference_qr(pk, orc):
    """prior.rs:97-110 at a size where the blocked device Cholesky runs several panels (d = 150: 5 panels of 32)."""
    n, d, k = 2500, 150, 6
    X, C0, mu0, s0 = _case(n, d, k, 0.25, seed=31, empty_dims=(d - 1,))
    ds = pk.Dataset(X)
    rng = np.random.default_rng(7)
    A = rng.standard_normal((d, d))
    cov = A @ A.T / d + 0.5 * np.eye(d)
    m0 = rng.standard_normal(d)
    prior = pk.Prior().with_mean_prior(m0, cov).with_isotropic_noise_prior(2.0, 1.5).with_transformation_precision(0.1)
    oprior = orc.Prior(mean=m0, mean_covariance=cov, isotropic_noise_alpha=2.0, isotropic_noise_beta=1.5,
                       transformation_precision=0.1)
    C, mu, s = C0, mu0, s0
    for it in range(3):
        new = pk.PPCAModel(s, C, mu).iterate_with_prior(ds, prior)
        (C, mu, s), (Cs, mus, ss) = both(orc, orc.iterate, X, None, C, mu, s, oprior)
        assert_close(new.transform, C, Cs, "C")
        assert_close(new.mean, mu, mus, "mu")
        assert_close(new.isotropic_noise, s, ss, "sigma")


@pytest.mark.parametrize("n,d,k,chunk", [(5000, 90, 12, 1024), (3000, 200, 16, 0), (2048, 33, 7, 512)])
def test_packed_host_streaming_equals_plain_host_streaming(pk, orc, n, d, k, chunk):
    """The compact host format (observed values + offsets + mask words) must give the plain out-of-core path's bits."""
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=17, empty_rows=(3,), empty_dims=(d - 1,))
    w = np.random.default_rng(2).random(n) + 0.5
    ctx = pk.get_context()
    ctx.set_chunk(chunk)
    try:
        model = pk.PPCAModel(s0, C0, mu0)
        a, llk_a = model._iterate(pk.HostDataset(X, w, pin=False), None)
        b, llk_b = model._iterate(pk.HostDataset(X, w, pin=True, packed=True), None)
    finally:
        ctx.set_chunk(0)
    assert np.array_equal(a.transform, b.transform) and np.array_equal(a.mean, b.mean)
    assert a.isotropic_noise == b.isotropic_noise and llk_a == llk_b
    (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C0, mu0, s0)
    assert_close(b.transform, Cw, Cs, "C")
    assert_close(b.mean, muw, mus, "mu")


# Every per-sample solve kernel and its padding rules: k <= 8 / <= 16 / <= 32 (lane owns a row), 33..48 / 49..64
# (register-tiled sweep, state padded to 48 / 64 with whole pivot blocks past k skipped), > 64 (generic).
@pytest.mark.parametrize("k", [3, 8, 9, 16, 17, 20, 24, 25, 31, 32, 33, 40, 41, 47, 48, 49, 55, 56, 57, 63, 64, 65, 70])
def test_state_size_sweep(pk, orc, k):
    n, d = 260, 96
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=k, empty_rows=(3,))
    w = np.random.default_rng(k).random(n) + 0.5
    ds = pk.Dataset(X, w)
    ctx = pk.get_context()
    before = ctx.variant_counts()
    model = pk.PPCAModel(0.6, C0, mu0)
    assert rel_err(model.llks(ds), orc.llks(X, C0, mu0, 0.6)) < TOL
    inf = model.infer(ds)
    (Z, COV), (Zs, COVs) = both(orc, orc.infer, X, C0, mu0, 0.6)
    assert_close(inf.states(), Z, Zs, "states")
    assert_close(np.stack(inf.covariances()), COV, COVs, "covariances")
    assert np.array_equal(inf.covariances()[3], np.eye(k))
    new, llk = model._iterate(ds, None)
    (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C0, mu0, 0.6)
    assert abs(llk - orc.llk(X, w, C0, mu0, 0.6)) <= TOL * abs(llk)
    assert_close(new.transform, Cw, Cs, "C")
    assert_close(new.mean, muw, mus, "mu")
    assert_close(new.isotropic_noise ** 2, sw ** 2, ss ** 2, "sigma^2")
    after = ctx.variant_counts()
    ran = {name for name, v in after.items() if name.startswith("solve") and v > before.get(name, 0)}
    want = ("solve_reg8" if k <= 8 else "solve_reg16" if k <= 16 else "solve_reg32" if k <= 32 else
            "solve_tile" if k <= 64 else "solve_generic")
    assert ran == {want}, (ran, want)


@pytest.mark.parametrize("k", [24, 48, 64])
def test_tiled_solve_on_ill_conditioned_systems(pk, orc, guard_ctx, k):
    """One feature in other units makes M_n = small + big v v^T: the register-tiled sweep has to keep the accuracy of an
    LU / Cholesky evaluation there (tools/solve_accuracy.py), on the FP64 contraction so only the solve is under test."""
    import dense_ref
    n, d = 400, 60
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=100 + k)
    X[:, 7] *= 1e3
    C0 = C0.copy()
    C0[7] *= 1e3
    ds = pk.Dataset(X)
    guard_ctx.set_gemm("dmma")
    model = pk.PPCAModel(0.5, C0, mu0)
    got = model.llks(ds)
    want = orc.llks(X, C0, mu0, 0.5)
    dense = np.array([dense_ref.llk_one(X[i], C0, mu0, 0.5) for i in range(n)])
    floor = _elementwise(dense, want)
    print(f"k {k}: tiled solve {_elementwise(got, want):.2e} floor {floor:.2e}")
    assert _elementwise(got, want) < TOL + 4.0 * floor


def test_dataset_from_device_is_the_host_dataset(pk, orc):
    """zero-copy ingestion (ppca_b200_dataset_from_device) of a torch tensor through
    __cuda_array_interface__ / DLPack gives the same masks, values, weights and EM step as Dataset(ndarray)."""
    import torch
    n, d, k = 3000, 70, 10
    X, C0, mu0, s0 = _case(n, d, k, 0.25, seed=21, empty_rows=(5,))
    X[17, 3] = np.inf
    X[18, 4] = -np.inf                                   # +-inf are missing values too (dataset.rs:19-22)
    w = np.random.default_rng(5).random(n) + 0.5
    ref = pk.Dataset(X, w)
    t = torch.from_numpy(X).cuda()
    tw = torch.from_numpy(w).cuda()
    ds = pk.Dataset.from_device(t, tw)
    want = ref.numpy()
    assert np.array_equal(np.isnan(ds.numpy()), np.isnan(want))
    assert np.array_equal(np.nan_to_num(ds.numpy()), np.nan_to_num(want))
    assert np.array_equal(ds.weights(), w)
    assert ds.empty_dimensions() == ref.empty_dimensions()
    # strided rows (a column slice of a wider tensor) and a DLPack-only producer
    wide = torch.full((n, d + 9), float("nan"), dtype=torch.float64, device="cuda")
    wide[:, 4:4 + d] = t
    ds2 = pk.Dataset.from_device(wide[:, 4:4 + d])
    assert np.array_equal(np.nan_to_num(ds2.numpy()), np.nan_to_num(want))

    class OnlyDLPack:
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, stream=None):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    ds3 = pk.Dataset.from_device(OnlyDLPack(t))
    assert np.array_equal(np.nan_to_num(ds3.numpy()), np.nan_to_num(want))
    # the export side: numpy() without leaving the device
    back = ds.to_torch()
    assert back.is_cuda and back.dtype == torch.float64
    assert np.array_equal(np.nan_to_num(back.cpu().numpy()), np.nan_to_num(want))
    assert np.array_equal(np.isnan(back.cpu().numpy()), np.isnan(want))
    # and the EM step is the host dataset's, bit for bit
    model = pk.PPCAModel(s0, C0, mu0)
    a, la = model._iterate(ref, None)
    b, lb = model._iterate(ds, None)
    assert la == lb and np.array_equal(a.transform, b.transform) and a.isotropic_noise == b.isotropic_noise
    (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, w, C0, mu0, s0)
    assert_close(b.transform, Cw, Cs, "C")
    # errors: wrong dtype, host tensor, non-positive weights reach mixture EM as in the host path
    with pytest.raises(TypeError):
        pk.Dataset.from_device(t.float())
    with pytest.raises(Exception):
        pk.Dataset.from_device(torch.from_numpy(X))
    bad = tw.clone()
    bad[7] = 0.0
    mix = pk.PPCAMix([pk.PPCAModel(1.0, C0, mu0), pk.PPCAModel(1.0, C0[:, ::-1].copy(), mu0)], np.zeros(2))
    with pytest.raises(Exception):
        mix.iterate(pk.Dataset.from_device(t, bad))


def test_mixture_and_posterior_samplers_on_device_have_the_reference_distributions(pk):
    """SURVEY §8(f3).  PPCAMix.sample (mix.rs:124-134): component ~ exp(log_weights), x | j ~ N(mu_j, C_j C_j^T + s_j^2 I),
    entries masked i.i.d.  PosteriorSampler (ppca_model.rs:595-626): x ~ N(C z + mu, C Sigma C^T + sigma^2 I).
    PosteriorSamplerMix (mix.rs:505-532): component ~ posterior of the row.  The reference is unseeded, so parity is
    distributional; a seed reproduces (counter-based RNG)."""
    rng = np.random.default_rng(9)
    d, n = 10, 300_000
    ks = (2, 4, 3)
    models = [pk.PPCAModel(0.3 + 0.2 * j, rng.standard_normal((d, k)), 4.0 * j + rng.standard_normal(d))
              for j, k in enumerate(ks)]
    wts = np.array([0.2, 0.5, 0.3])
    mix = pk.PPCAMix(models, np.log(wts))
    a = mix.sample(n, 0.25, seed=3).numpy()
    fin = np.isfinite(a)
    assert a.shape == (n, d) and abs(fin.mean() - 0.75) < 5e-3
    # the mixture mean and second moment of the draw
    mean_true = sum(w * m.mean.reshape(-1) for w, m in zip(wts, models))
    mean = np.nanmean(a, axis=0)
    sec_true = sum(w * (m.transform @ m.transform.T + m.isotropic_noise ** 2 * np.eye(d)
                        + np.outer(m.mean.reshape(-1), m.mean.reshape(-1))) for w, m in zip(wts, models))
    x0 = np.where(fin, a, 0.0)
    pair = fin.astype(np.float64).T @ fin.astype(np.float64)
    sec = (x0.T @ x0) / pair
    assert np.max(np.abs(mean - mean_true)) < 0.05
    assert np.max(np.abs(sec - sec_true)) < 0.02 * np.max(np.abs(sec_true))
    # component frequencies through the posterior of the fully observed draws (means are 4 apart: well separated)
    full = mix.sample(50_000, 0.0, seed=4)
    freq = np.bincount(np.argmax(mix.infer_cluster(full), axis=1), minlength=3) / 50_000
    assert np.max(np.abs(freq - wts)) < 0.01
    assert np.array_equal(mix.sample(500, 0.25, seed=3).numpy(), a[:500], equal_nan=True)
    assert len(mix.sample(0, 0.1)) == 0
    with pytest.raises(ValueError):
        mix.sample(10, 1.5)

    # posterior sampler of one model: many draws of the same inferred sample
    model = models[1]
    x = model.sample(1, 0.4, seed=8).numpy()
    reps = 200_000
    inf = model.infer(pk.Dataset(np.repeat(x, reps, axis=0)))
    z, S = inf.states()[0], inf.covariances()[0]
    draw = inf.posterior_sampler().sample(seed=5).numpy()
    assert draw.shape == (reps, d) and np.isfinite(draw).all()          # posterior samples are fully observed
    Cm, mu, s = model.transform, model.mean.reshape(-1), model.isotropic_noise
    assert np.max(np.abs(draw.mean(axis=0) - (Cm @ z + mu))) < 6 * np.sqrt(np.max(np.diag(Cm @ S @ Cm.T) + s * s) / reps)
    cov_true = Cm @ S @ Cm.T + s * s * np.eye(d)
    assert np.max(np.abs(np.cov(draw.T) - cov_true)) < 0.03 * np.max(np.abs(cov_true))
    # mixture posterior sampler: rows whose posterior is concentrated on one component are drawn from that component
    data = mix.sample(20_000, 0.3, seed=6)
    infm = mix.infer(data)
    post = infm.posteriors()
    dm = infm.posterior_sampler().sample(seed=7).numpy()
    assert dm.shape == (20_000, d) and np.isfinite(dm).all()
    sm = mix.smooth(data).numpy()
    sure = post.max(axis=1) > 0.999
    assert sure.mean() > 0.9
    resid = dm[sure] - sm[sure]                       # draw = smoothed + C L xi + sigma eps around the same component
    assert abs(resid.mean()) < 0.02
    assert np.array_equal(infm.posterior_sampler().sample(seed=7).numpy(), dm)
    # a covariance that is not positive definite raises, as the reference's expect() (ppca_model.rs:582-586)
    bad = pk.InferredMasked(np.zeros((4, 2)), np.stack([np.array([[1.0, 2.0], [2.0, 1.0]])] * 4), models[0])
    with pytest.raises(Exception):
        bad.posterior_sampler().sample()


@pytest.mark.parametrize("n,d,k", [(60, 37, 5), (40, 100, 16), (9, 70, 33)])
def test_full_covariances_on_device(pk, orc, n, d, k):
    """InferredMasked.smoothed_covariances / extrapolated_covariances (ppca_model.rs:471-477, 517-534) and the
    mixture forms (mix.rs:422-437, 466-481): d x d matrices made on the device (ppca_b200_covariance_full) against the
    oracle's restatement."""
    X, C0, mu0, s0 = _case(n, d, k, 0.3, seed=61)
    X[3, :] = np.nan                       # nothing observed
    X[4, :] = 1.5                          # nothing missing: the extrapolated covariance is all zeros (:520-522)
    model = pk.PPCAModel(0.4, C0, mu0)
    ds = pk.Dataset(X)
    inf = model.infer(ds)
    covs = inf.covariances()
    sm = inf.smoothed_covariances(model)
    ex = inf.extrapolated_covariances(model, ds)
    assert len(sm) == n and sm[0].shape == (d, d)
    for i in range(n):
        assert rel_err(sm[i], orc.smoothed_covariance(C0, 0.4, covs[i])) < 1e-12
        want = orc.extrapolated_covariance(C0, 0.4, covs[i], X[i])
        assert np.array_equal(ex[i] == 0.0, want == 0.0) or rel_err(ex[i], want) < 1e-12
        assert np.max(np.abs(ex[i] - want)) <= 1e-12 * max(np.max(np.abs(want)), 1.0)
    assert not ex[4].any() and np.array_equal(ex[3], sm[3])
    assert rel_err(np.stack([np.diag(m) for m in sm]), inf.smoothed_covariances_diagonal(model).numpy()) < 1e-12
    other = pk.PPCAModel(0.7, C0[:, ::-1].copy(), mu0 + 0.3)
    mix = pk.PPCAMix([model, other], np.log([0.4, 0.6]))
    im = mix.infer(ds)
    post = im.posteriors()
    msm = im.smoothed_covariances(mix)
    mex = im.extrapolated_covariances(mix, ds)
    comps = [model, other]
    for i in range(0, n, 7):
        cj = [orc.smoothed_covariance(m.transform, m.isotropic_noise, im._inf[j].covariances()[i]) for j, m in enumerate(comps)]
        smj = [im._inf[j].states()[i] @ m.transform.T + m.mean.reshape(-1) for j, m in enumerate(comps)]
        exj = [np.where(np.isfinite(X[i]), X[i], v) for v in smj]
        assert rel_err(msm[i], orc.mix_covariance(post[i], smj, cj)) < 1e-11
        assert rel_err(mex[i], orc.mix_covariance(post[i], exj, cj)) < 1e-11


@pytest.mark.parametrize("chunk", [0, 1024])
def test_generated_dataset_equals_the_stored_rows(pk, orc, chunk):
    """ppca_b200_iterate_generated (chunks regenerated on the device, never stored: the out-of-core form of BASELINE
    configs[2]) gives the EM step of the same rows held as a resident Dataset, for a row range that starts past row 0."""
    n, d, k_true, k = 5000, 70, 6, 5
    ctx = pk.get_context()
    gen = pk.GeneratedDataset(n, d, k_true, 0.1, 0.3, seed=99, row_begin=1234)
    stored = gen.materialize()
    assert len(stored) == n
    whole = pk.Dataset.synthetic(1234 + n, d, k_true, 0.1, 0.3, seed=99).numpy()
    assert np.array_equal(stored.numpy(), whole[1234:], equal_nan=True)
    mixed = pk.Dataset.synthetic(1234 + 700, d, k_true, 0.1, 0.3, n_components=3, seed=98).numpy()
    part = pk.Dataset.synthetic(700, d, k_true, 0.1, 0.3, n_components=3, seed=98, row_begin=1234).numpy()
    assert np.array_equal(part, mixed[1234:], equal_nan=True)     # row ranges of ONE dataset, mixtures of truths too
    _, C0, mu0, s0 = _case(50, d, k, 0.2, seed=3)
    model = pk.PPCAModel(s0, C0, mu0)
    ctx.set_chunk(chunk)
    try:
        a, la = model._iterate(gen, None)
    finally:
        ctx.set_chunk(0)
    b, lb = model._iterate(stored, None)
    assert abs(la - lb) <= 1e-12 * abs(lb)
    assert rel_err(a.transform, b.transform) < 1e-11 and rel_err(a.mean, b.mean) < 1e-11
    assert abs(a.isotropic_noise - b.isotropic_noise) < 1e-12 * b.isotropic_noise
    X = stored.numpy()
    (Cw, muw, sw), (Cs, mus, ss) = both(orc, orc.iterate, X, None, C0, mu0, s0)
    assert_close(a.transform, Cw, Cs, "C")
    assert_close(a.isotropic_noise ** 2, sw ** 2, ss ** 2, "sigma^2")
    assert a.iterate(gen).llk(stored) > model.llk(stored)


def test_mixture_chunk_loop_replays_from_a_cuda_graph(pk, orc):
    """The mixture EM chunk loop (~30 launches per component and chunk) is captured into a CUDA graph the second time a
    (dataset, shapes, buffers) key is seen and replayed afterwards; the replays must give the eager pass's numbers bit for
    bit, follow a CHANGED model (sigma is read from device memory, not baked into the launches), and match the oracle."""
    n, d = 3000, 24
    X, base, logw = _mix_case(pk, n, d, (3, 5, 2))
    w = np.random.default_rng(1).random(X.shape[0]) + 0.5
    ds = pk.Dataset(X, w)

    def make(scale):
        return pk.PPCAMix([pk.PPCAModel(scale * s, C * (2.0 - scale) if scale != 1.0 else C, mu) for C, mu, s in base], logw)

    ctx = pk.get_context()
    ctx.set_chunk(1024)                       # three chunks: the running-maximum rescaling is inside the graph too
    try:
        mixA, mixB = make(1.0), make(1.7)
        before = ctx.variant_counts()["graph_replays"]
        outs = [mixA._iterate(ds, None) for _ in range(4)]          # eager, captured, replayed, replayed
        replays = ctx.variant_counts()["graph_replays"] - before
        assert replays >= 2, replays
        for o, l in outs[1:]:
            assert l == outs[0][1]
            for a, b in zip(o.models, outs[0][0].models):
                assert np.array_equal(a.transform, b.transform) and np.array_equal(a.mean, b.mean)
                assert a.isotropic_noise == b.isotropic_noise
            assert np.array_equal(o.log_weights, outs[0][0].log_weights)
        got, llk = mixB._iterate(ds, None)                          # same key, different model: replayed again
        assert ctx.variant_counts()["graph_replays"] - before == replays + 1
    finally:
        ctx.set_chunk(0)
    models = [(m.transform, m.mean.reshape(-1), m.isotropic_noise) for m in mixB.models]
    with orc.stable():
        new_models, new_lw = orc.mix_iterate(X, w, models, mixB.log_weights)
    assert abs(llk - orc.mix_llk(X, w, models, mixB.log_weights)) <= TOL * abs(llk)
    for j, (Cw, muw, sw) in enumerate(new_models):
        assert rel_err(got.models[j].transform, Cw) < TOL
        assert rel_err(got.models[j].mean.reshape(-1), muw) < TOL
        assert abs(got.models[j].isotropic_noise - sw) < TOL * sw
    assert rel_err(got.log_weights, new_lw) < TOL
