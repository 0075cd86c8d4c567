"""CPU tests (no GPU): C ABI surface, host-side mirror logic, serialisation, sharding plumbing over gloo."""
import ctypes
import os
import re
import socket
import struct

import numpy as np
import pytest

from helpers import init_model, make_data, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- the C ABI -----------------------------------------------------------------------------------
def _header_symbols():
    text = open(os.path.join(ROOT, "include", "ppca_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ppca_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from ppca_rs_b200 import _native as nat
    lib = ctypes.CDLL(nat.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 38
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ppca_b200.h but not exported"
    assert sorted(nat._PROTOTYPES) == syms            # the ctypes binding covers the whole header
    assert nat.lib().ppca_b200_abi_version() == 1


def test_stats_len_matches_python_layout():
    from ppca_rs_b200 import _native as nat
    from ppca_rs_b200.distributed import stats_len
    for d, k in [(3, 2), (200, 16), (2048, 64), (37, 5), (512, 32), (1024, 48)]:
        assert nat.lib().ppca_b200_em_stats_len(d, k) == stats_len(d, k)
    assert nat.lib().ppca_b200_em_stats_len(0, 4) == 0


def test_no_cpu_fallback():
    """Without a device every compute entry point fails loudly (never routes to a CPU path)."""
    import ppca_rs_b200 as pk
    if pk.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pk.NativeError):
        pk.Dataset(np.zeros((4, 3)))
    with pytest.raises(pk.NativeError):
        pk.Context(0)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ppca_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("test oracle", ""), f


# ---- host-side model logic ------------------------------------------------------------------------
def test_model_constructor_and_getters():
    import ppca_rs_b200 as pk
    C = np.arange(6.0).reshape(3, 2)
    m = pk.PPCAModel(isotropic_noise=0.1, transform=C, mean=np.array([[0.0], [1.0], [0.0]]))  # examples/toy_model.py
    assert m.output_size == 3 and m.state_size == 2 and m.n_parameters == 1 + 6 + 3   # ppca_model.rs:107-109
    assert m.mean.shape == (3,) and np.array_equal(m.mean, [0, 1, 0])
    assert np.array_equal(pk.PPCAModel(0.1, C, np.array([[0.0, 1.0, 0.0]])).mean, [0, 1, 0])  # 1 x n accepted
    with pytest.raises(ValueError):
        pk.PPCAModel(0.1, C, np.zeros((3, 3)))       # src/utils.rs:17-21 panics on a non-vector
    assert np.allclose(m.singular_values, np.sqrt(np.linalg.norm(C, axis=0)))  # ppca_model.rs:113-121 quirk
    assert "PPCAModel(isotropic_noise=0.1" in repr(m)


def test_to_canonical_matches_oracle_and_reference_rules():
    import ppca_rs_b200 as pk
    from oracle import oracle as orc
    rng = np.random.default_rng(0)
    C = rng.standard_normal((30, 5))
    canon = pk.PPCAModel(0.3, C, np.zeros(30)).to_canonical()
    assert rel_err(canon.transform, orc.to_canonical(C)) < 1e-10
    assert np.all(canon.transform.sum(axis=0) >= 0)
    assert canon.isotropic_noise == 0.3


def test_prior_validation():
    import ppca_rs_b200 as pk
    p = pk.Prior()
    with pytest.raises(ValueError):
        p.with_isotropic_noise_prior(-1.0, 1.0)      # prior.rs:50
    with pytest.raises(ValueError):
        p.with_transformation_precision(-0.1)        # prior.rs:61
    with pytest.raises(ValueError):
        p.with_mean_prior(np.zeros(3), np.eye(4))    # prior.rs:33-34
    q = p.with_mean_prior(np.zeros((1, 3)), 2.0 * np.eye(3)).with_transformation_precision(0.5)
    assert p.mean is None and q.transformation_precision == 0.5   # builder returns new objects
    assert np.allclose(q.mean_precision, 0.5 * np.eye(3))


def test_mix_constructor():
    import ppca_rs_b200 as pk
    m1 = pk.PPCAModel(1.0, np.ones((4, 2)), np.zeros(4))
    m2 = pk.PPCAModel(1.0, np.ones((4, 3)), np.zeros(4))
    mix = pk.PPCAMix([m1, m2], np.log([1.0, 3.0]))
    assert np.allclose(mix.weights, [0.25, 0.75])                 # mix.rs:69 log-softmax normalisation
    assert mix.state_sizes == [2, 3] and mix.output_size == 4
    assert mix.n_parameters == m1.n_parameters + m2.n_parameters + 1   # mix.rs:96-104
    with pytest.raises(ValueError):
        pk.PPCAMix([], [])
    with pytest.raises(ValueError):
        pk.PPCAMix([m1, pk.PPCAModel(1.0, np.ones((5, 2)), np.zeros(5))], [0.0, 0.0])


# ---- bincode layouts (src/python_bindings.rs:66-79,388-401,571-584) --------------------------------
def test_bincode_model_layout_and_roundtrip():
    import pickle
    import ppca_rs_b200 as pk
    C = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    m = pk.PPCAModel(0.5, C, np.array([7.0, 8.0, 9.0]))
    raw = m.dump()
    # f64 sigma | DMatrix: u64 len, column-major data, u64 nrows, u64 ncols | DVector: u64 len, data, u64 nrows
    expect = struct.pack("<d", 0.5) + struct.pack("<Q", 6) + struct.pack("<6d", 1, 3, 5, 2, 4, 6) + struct.pack("<QQ", 3, 2)
    expect += struct.pack("<Q", 3) + struct.pack("<3d", 7, 8, 9) + struct.pack("<Q", 3)
    assert raw == expect
    back = pk.PPCAModel.load(raw)
    assert np.array_equal(back.transform, C) and back.isotropic_noise == 0.5 and np.array_equal(back.mean, m.mean)
    assert np.array_equal(pickle.loads(pickle.dumps(m)).transform, C)
    with pytest.raises(Exception):
        pk.PPCAModel.load(raw[:-3])
    mix = pk.PPCAMix([m, m], [0.0, 0.0])
    again = pk.PPCAMix.load(mix.dump())
    assert np.allclose(again.log_weights, mix.log_weights) and np.array_equal(again.models[1].transform, C)
    assert mix.dump()[:16] == struct.pack("<QQ", 3, 2)            # output_size, number of models


def test_bincode_dataset_layout():
    from ppca_rs_b200 import bincode
    X = np.array([[1.0, np.nan, 3.0], [np.nan, np.nan, 6.0]])
    raw = bincode.dump_dataset(X, np.array([1.0, 2.0]))
    # Vec len | per sample: DVector (len, data, nrows) + BitVec (n blocks, u32 blocks LSB-first, nbits) | weights
    assert raw[:8] == struct.pack("<Q", 2)
    first_mask = raw[8 + 8 + 24 + 8:8 + 8 + 24 + 8 + 8 + 4 + 8]
    assert first_mask == struct.pack("<Q", 1) + struct.pack("<I", 0b101) + struct.pack("<Q", 3)
    x2, w2 = bincode.load_dataset(raw)
    assert np.array_equal(np.isnan(x2), np.isnan(X)) and np.array_equal(x2[np.isfinite(X)], X[np.isfinite(X)])
    assert np.array_equal(w2, [1.0, 2.0])


# ---- sharding plumbing ------------------------------------------------------------------------------
def test_shard_bounds_cover_everything():
    from ppca_rs_b200.distributed import shard_bounds
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            bounds = [shard_bounds(n, world, r) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, X, w, C0, mu0, s0, iters, out):
    import torch.distributed as dist
    from numpy_engine import HostShard, NumpyEngine
    from ppca_rs_b200.distributed import ShardedPPCA, shard_bounds
    from ppca_rs_b200.model import PPCAModel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_bounds(X.shape[0], world, rank)
    state = ShardedPPCA(None, HostShard(X[lo:hi], w[lo:hi]), PPCAModel(s0, C0, mu0), group=dist, engine=NumpyEngine())
    llks = [state.step() for _ in range(iters)]
    if rank == 0:
        np.savez(out, C=state.model.transform, mu=state.model.mean, s=state.model.isotropic_noise, llk=np.array(llks))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_em_world_size_2_gloo(tmp_path):
    """Two gloo ranks, each with half of the rows: local statistics -> all-reduce(SUM) -> replicated finish must
    equal the single-process EM on the whole dataset (every statistic is additive over samples)."""
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    n, d, k, iters = 301, 14, 3, 3
    X = make_data(n, d, k, 0.25, seed=12, empty_rows=(5,))
    w = np.random.default_rng(1).random(n) + 0.5
    C0, mu0, s0 = init_model(d, k)
    out = str(tmp_path / "rank0.npz")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    mp.spawn(_gloo_worker, args=(2, _free_port(), X, w, C0, mu0, s0, iters, out), nprocs=2, join=True)
    got = np.load(out)
    C, mu, s = C0, mu0, s0
    llks = []
    with orc.stable():
        for _ in range(iters):
            llks.append(orc.llk(X, w, C, mu, s))
            C, mu, s = orc.iterate(X, w, C, mu, s)
    assert rel_err(got["C"], C) < 1e-9 and rel_err(got["mu"], mu) < 1e-9
    assert abs(float(got["s"]) - s) < 1e-9 * s
    assert rel_err(got["llk"], np.array(llks)) < 1e-10


# ---- DataFrameAdapter (python/ppca_rs/__init__.py:121-433): host logic, no GPU (dataset_factory injected) ----------
class _FakeDataset:
    def __init__(self, x):
        self.x = np.array(x, dtype=np.float64)

    def numpy(self):
        return self.x


def _adapter_frame(seed=0, n_keys=40, n_dims=7):
    import pandas as pd
    rng = np.random.default_rng(seed)
    rows = []
    for day in range(n_keys // 4):
        for shop in "abcd":
            for dim in range(n_dims):
                if rng.random() < 0.3:
                    continue                                   # a missing combination -> NaN in the dataset
                rows.append({"day": day, "shop": shop, "product": f"p{dim % 3}", "size": dim // 3,
                             "sales": float(rng.standard_normal())})
    df = pd.DataFrame(rows)
    return df.sample(frac=1.0, random_state=1).reset_index(drop=True)   # row order must not matter


def test_dataframe_adapter_matches_reference_semantics():
    pytest.importorskip("pandas")
    from ppca_rs_b200.adapter import DataFrameAdapter, DataFrameAdapterDescription
    df = _adapter_frame()
    ad = DataFrameAdapter.from_pandas(df, keys=["day", "shop"], dimensions=["product", "size"], metric="sales",
                                      dataset_factory=_FakeDataset)
    # the reference's construction, restated: sorted distinct dimension tuples, groupby(keys) order, scatter per group
    dims = sorted(set(zip(df["product"], df["size"])))
    assert [tuple(r) for r in ad.dimension_idx[["product", "size"]].itertuples(index=False)] == dims
    assert list(ad.dimension_idx["__dim_idx"]) == list(range(len(dims)))
    groups = sorted(set(zip(df["day"], df["shop"])))
    assert [tuple(r) for r in ad.sample_idx[["day", "shop"]].itertuples(index=False)] == groups
    want = np.full((len(groups), len(dims)), np.nan)
    for _, row in df.iterrows():
        want[groups.index((row["day"], row["shop"])), dims.index((row["product"], row["size"]))] = row["sales"]
    got = ad.dataset.numpy()
    assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])

    # description round trip (to_json / from_json) re-adapts new data onto the SAME dimension numbering
    desc = DataFrameAdapterDescription.from_json(ad.description().to_json())
    assert desc.dimension_idx == [list(t) for t in dims] and desc.keys == ["day", "shop"] and desc.metric == "sales"
    sub = df[df["day"] < 3]
    ad2 = desc.adapt_pandas(sub, dataset_factory=_FakeDataset)
    assert ad2.dimensions == ["product", "size"] and ad2.dataset.numpy().shape[1] == len(dims)
    assert np.allclose(ad2.dataset.numpy(), want[: len(set(zip(sub["day"], sub["shop"])))], equal_nan=True)

    # back to the long format: every (sample, dimension) pair, keys and dimensions attached
    long = ad.convert_datasets({"sales_hat": _FakeDataset(np.nan_to_num(want, nan=-1.0)), "raw": ad.dataset})
    assert list(long.columns) == ["day", "shop", "product", "size", "sales_hat", "raw"]
    assert len(long) == want.size
    merged = long.merge(df, on=["day", "shop", "product", "size"], how="left")
    obs = ~merged["sales"].isna()
    assert np.allclose(merged.loc[obs, "sales_hat"], merged.loc[obs, "sales"])
    assert np.all(merged.loc[~obs, "sales_hat"] == -1.0) and merged.loc[~obs, "raw"].isna().all()
    one = ad.convert_dataset(ad.dataset, column_name="x")
    assert list(one.columns) == ["day", "shop", "product", "size", "x"]


def test_dataframe_adapter_unknown_dimensions_are_dropped():
    pytest.importorskip("pandas")
    import pandas as pd
    from ppca_rs_b200.adapter import DataFrameAdapter
    df = pd.DataFrame({"k": [0, 0, 1, 1], "dim": ["a", "b", "a", "zzz"], "v": [1.0, 2.0, 3.0, 4.0]})
    idx = pd.DataFrame({"__dim_idx": [0, 1], "dim": ["a", "b"]})
    ad = DataFrameAdapter.from_pandas(df, keys=["k"], dimension_idx=idx, metric="v", dataset_factory=_FakeDataset)
    assert ad.dimensions == ["dim"]
    assert np.allclose(ad.dataset.numpy(), [[1.0, 2.0], [3.0, np.nan]], equal_nan=True)


def test_rust_ffi_declarations_are_in_sync_with_the_header(tmp_path):
    """bindings/rust/ppca_b200_sys.rs is generated from include/ppca_b200.h (tools/gen_rust_ffi.py): one `pub fn` per
    exported symbol, regenerating reproduces the committed file."""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    committed = open(os.path.join(root, "bindings", "rust", "ppca_b200_sys.rs")).read()
    header = open(os.path.join(root, "include", "ppca_b200.h")).read()
    declared = set(re.findall(r"\b(ppca_b200_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", header, flags=re.S)))
    bound = set(re.findall(r"pub fn (ppca_b200_\w+)\(", committed))
    assert declared == bound
    subprocess.check_call([sys.executable, os.path.join(root, "tools", "gen_rust_ffi.py")], stdout=subprocess.DEVNULL)
    assert open(os.path.join(root, "bindings", "rust", "ppca_b200_sys.rs")).read() == committed


def test_compat_alias_exposes_the_reference_module_names():
    """compat.install_as_ppca_rs(): `from ppca_rs import ...` and `from ppca_rs.ppca_rs import ...` resolve to this package
    (names of python/ppca_rs/__init__.py and of the pyO3 module, src/python_bindings.rs:15-26)."""
    import importlib
    import sys
    import ppca_rs_b200 as pk
    import ppca_rs_b200.compat as compat
    saved = {k: sys.modules.pop(k) for k in ("ppca_rs", "ppca_rs.ppca_rs") if k in sys.modules}
    try:
        compat.install_as_ppca_rs()
        top = importlib.import_module("ppca_rs")
        native = importlib.import_module("ppca_rs.ppca_rs")
        for name in ("Dataset", "PPCAModel", "PPCAMix", "Prior", "PPCATrainer", "PPCAMixTrainer", "DataFrameAdapter",
                     "DataFrameAdapterDescription", "InferredMasked", "InferredMaskedMix"):
            assert getattr(top, name) is getattr(pk, name)
        for name in ("Dataset", "Prior", "PPCAModel", "PPCAMix", "InferredMasked", "InferredMaskedMix"):
            assert getattr(native, name) is getattr(pk, name)
        compat.install_as_ppca_rs()                     # idempotent on its own alias
        sys.modules["ppca_rs"] = type(sys)("ppca_rs")   # a "real" module already imported: refuse to shadow it
        with pytest.raises(ImportError):
            compat.install_as_ppca_rs()
        compat.install_as_ppca_rs(force=True)
    finally:
        for k in ("ppca_rs", "ppca_rs.ppca_rs"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)


def test_posterior_samplers_carry_the_model_without_a_gpu():
    """InferredMasked / InferredMaskedMix keep the model they came from (ppca_model.rs:428-432, mix.rs:357-362), so
    posterior_sampler().sample() needs no argument; checked here on the host objects alone."""
    from ppca_rs_b200.model import InferredMasked, InferredMaskedMix, PPCAMix, PPCAModel
    rng = np.random.default_rng(0)
    model = PPCAModel(0.5, rng.standard_normal((6, 2)), np.zeros(6))
    inf = InferredMasked(rng.standard_normal((5, 2)), np.stack([np.eye(2)] * 5), model)
    smp = inf.posterior_sampler()
    assert smp._model is model
    mix = PPCAMix([model, model], np.log([0.5, 0.5]))
    infm = InferredMaskedMix(np.log(np.full((5, 2), 0.5)), [inf, inf], mix)
    assert infm.posterior_sampler()._mix is mix
    import inspect
    from ppca_rs_b200.model import PosteriorSampler, PosteriorSamplerMix
    for cls in (PosteriorSampler, PosteriorSamplerMix):
        params = list(inspect.signature(cls.sample).parameters.values())[1:]
        assert all(p.default is not inspect.Parameter.empty for p in params)   # callable with no arguments


def test_sharded_front_end_picks_the_native_collective_only_with_a_group():
    """ShardedPPCA: engine given (CPU tests) -> spelled-out torch protocol; no group -> single process."""
    from numpy_engine import HostShard, NumpyEngine
    from ppca_rs_b200.distributed import ShardedPPCA
    from ppca_rs_b200.model import PPCAModel
    X = make_data(50, 6, 2, 0.2, seed=1)
    C0, mu0, s0 = init_model(6, 2)
    st = ShardedPPCA(None, HostShard(X, np.ones(50)), PPCAModel(s0, C0, mu0), group=None, engine=NumpyEngine())
    assert st.native is False
    assert ShardedPPCA(None, HostShard(X, np.ones(50)), PPCAModel(s0, C0, mu0), group=None).native is True
    llk = st.step()
    assert np.isfinite(llk)


def test_pack_host_builds_the_compact_format_without_a_gpu():
    """ppca_b200_pack_host (host-only): observed values row-major, row offsets, bit-vec style mask words."""
    import ctypes as C
    from ppca_rs_b200 import _native as nat
    rng = np.random.default_rng(3)
    n, d = 5000, 70
    X = rng.standard_normal((n, d))
    X[rng.random((n, d)) < 0.3] = np.nan
    X[7] = np.nan
    X[11, 3] = np.inf
    dw = (d + 31) // 32
    rowptr = np.empty(n + 1, dtype=np.int64)
    p64, p32 = C.POINTER(C.c_int64), C.POINTER(C.c_uint32)
    nat.check(nat.lib().ppca_b200_pack_host(nat.dptr(X), n, d, None, rowptr.ctypes.data_as(p64), None))
    fin = np.isfinite(X)
    assert rowptr[0] == 0 and np.array_equal(np.diff(rowptr), fin.sum(axis=1))
    vals = np.empty(int(rowptr[n]))
    maskw = np.empty((n, dw), dtype=np.uint32)
    nat.check(nat.lib().ppca_b200_pack_host(nat.dptr(X), n, d, nat.dptr(vals), rowptr.ctypes.data_as(p64),
                                            maskw.ctypes.data_as(p32)))
    assert np.array_equal(vals, X[fin])
    bits = np.zeros((n, dw * 32), dtype=bool)
    bits[:, :d] = fin
    want = np.packbits(bits, axis=1, bitorder="little").view(np.uint32).reshape(n, dw)
    assert np.array_equal(maskw, want)


def test_device_ingestion_front_end_validates_its_inputs():
    """Dataset.from_device (ppca_b200_dataset_from_device): the host-side checks of the __cuda_array_interface__ /
    DLPack view — dtype, rank, strides — run without a GPU."""
    from ppca_rs_b200 import model as M

    class Fake:
        def __init__(self, shape, typestr="<f8", strides=None, ptr=4096):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "strides": strides,
                                             "data": (ptr, False), "version": 3}

    ptr, shape, strides, keep = M._device_view(Fake((5, 3)), 2, "array")
    assert shape == (5, 3) and strides == (3, 1) and ptr.value == 4096
    _, _, strides, _ = M._device_view(Fake((5, 3), strides=(80, 8)), 2, "array")
    assert strides == (10, 1)                                     # bytes -> elements, padded rows
    with pytest.raises(TypeError):
        M._device_view(Fake((5, 3), typestr="<f4"), 2, "array")
    with pytest.raises(ValueError):
        M._device_view(Fake((5,)), 2, "array")
    with pytest.raises(ValueError):
        M._device_view(Fake((5, 3), strides=(20, 4)), 2, "array")
    with pytest.raises(TypeError):
        M._device_view(object(), 2, "array")


def test_generated_dataset_describes_a_row_range():
    import ppca_rs_b200 as pk
    if pk.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pk.NativeError):                           # no context without a device: never a CPU path
        pk.GeneratedDataset(10, 4, 2)
    with pytest.raises(ValueError):
        pk.GeneratedDataset(-1, 4, 2, ctx=object())


def test_bench_reference_arm_emits_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): ONE JSON line with the contract's keys,
    `impl`, a `cpu_baseline` describing the run and an `e2e` block that repeats the line's own value (no device copies)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [ln for ln in out.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert set(j["config"]) >= {"workload", "rows_per_gpu", "d", "k", "components"}
