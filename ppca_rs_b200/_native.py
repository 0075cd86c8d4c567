"""ctypes binding of libppca_b200.so (the C ABI declared in include/ppca_b200.h).

There is no CPU fallback: if the library is missing, or no sm_100a device is present, every compute call
raises.  Nothing in this package imports the test oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPCA_B200_LIB") or os.path.join(_HERE, "libppca_b200.so")  # override: A/B runs of two builds

c_ctx_p = C.c_void_p
c_ds_p = C.c_void_p
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)


class CPrior(C.Structure):
    """ppca_b200_prior (prior.rs:8-29)."""

    _fields_ = [
        ("has_mean_prior", C.c_int32),
        ("mean", c_dp),
        ("mean_precision", c_dp),
        ("has_isotropic_noise_prior", C.c_int32),
        ("isotropic_noise_alpha", C.c_double),
        ("isotropic_noise_beta", C.c_double),
        ("transformation_precision", C.c_double),
    ]


ERR_PRECISION = 6  # PPCA_ERR_PRECISION (include/ppca_b200.h)


class NativeError(RuntimeError):
    """Raised where the reference panics (pyO3 PanicException) or returns a PyException."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"ppca_b200 error {code}: {msg}")
        self.code = code


# name -> (restype, argtypes).  Must list every symbol of include/ppca_b200.h (checked by tests).
_PROTOTYPES = {
    "ppca_b200_abi_version": (C.c_int32, []),
    "ppca_b200_last_error": (C.c_char_p, []),
    "ppca_b200_device_count": (C.c_int32, [c_ip]),
    "ppca_b200_ctx_create": (C.c_int32, [C.c_int32, C.c_void_p, C.POINTER(c_ctx_p)]),
    "ppca_b200_ctx_destroy": (C.c_int32, [c_ctx_p]),
    "ppca_b200_ctx_synchronize": (C.c_int32, [c_ctx_p]),
    "ppca_b200_ctx_stream": (C.c_int32, [c_ctx_p, C.POINTER(C.c_void_p)]),
    "ppca_b200_ctx_set_chunk": (C.c_int32, [c_ctx_p, C.c_int64]),
    "ppca_b200_ctx_set_gemm": (C.c_int32, [c_ctx_p, C.c_int32, C.c_int32]),
    "ppca_b200_ctx_set_guard": (C.c_int32, [c_ctx_p, C.c_int32, C.c_int32]),
    "ppca_b200_ctx_launch_count": (C.c_int32, [c_ctx_p, C.POINTER(C.c_int64)]),
    "ppca_b200_ctx_variant_counts": (C.c_int32, [c_ctx_p, C.POINTER(C.c_int64)]),
    "ppca_b200_ctx_set_profiling": (C.c_int32, [c_ctx_p, C.c_int32]),
    "ppca_b200_ctx_last_profile": (C.c_int32, [c_ctx_p, c_dp]),
    "ppca_b200_dataset_from_host": (C.c_int32, [c_ctx_p, c_dp, C.c_int64, C.c_int32, c_dp, C.POINTER(c_ds_p)]),
    "ppca_b200_mix_sample": (
        C.c_int32, [c_ctx_p, C.c_int64, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_uint64,
                    C.POINTER(c_ds_p)]),
    "ppca_b200_posterior_sample": (
        C.c_int32, [c_ctx_p, C.c_int64, C.c_int32, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.POINTER(c_dp),
                    C.POINTER(c_dp), C.c_uint64, C.POINTER(c_ds_p)]),
    "ppca_b200_dataset_from_device": (
        C.c_int32, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(c_ds_p)]),
    "ppca_b200_dataset_to_device": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int64, C.c_int64, C.c_void_p]),
    "ppca_b200_dataset_synthetic": (
        C.c_int32,
        [c_ctx_p, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_uint64, C.POINTER(c_ds_p)],
    ),
    "ppca_b200_dataset_synthetic_rows": (
        C.c_int32,
        [c_ctx_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_uint64,
         C.POINTER(c_ds_p)],
    ),
    "ppca_b200_model_sample": (
        C.c_int32,
        [c_ctx_p, C.c_int64, C.c_int32, C.c_int32, c_dp, c_dp, C.c_double, C.c_double, C.c_uint64, C.POINTER(c_ds_p)],
    ),
    "ppca_b200_dataset_with_weights": (C.c_int32, [c_ctx_p, c_ds_p, c_dp, C.POINTER(c_ds_p)]),
    "ppca_b200_dataset_len": (C.c_int32, [c_ds_p, C.POINTER(C.c_int64)]),
    "ppca_b200_dataset_output_size": (C.c_int32, [c_ds_p, c_ip]),
    "ppca_b200_dataset_to_host": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int64, C.c_int64, c_dp]),
    "ppca_b200_dataset_weights": (C.c_int32, [c_ctx_p, c_ds_p, c_dp]),
    "ppca_b200_dataset_empty_dimensions": (C.c_int32, [c_ctx_p, c_ds_p, C.POINTER(C.c_uint8)]),
    "ppca_b200_dataset_slice": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int64, C.c_int64, C.POINTER(c_ds_p)]),
    "ppca_b200_dataset_concat": (C.c_int32, [c_ctx_p, C.POINTER(c_ds_p), C.c_int32, C.POINTER(c_ds_p)]),
    "ppca_b200_dataset_destroy": (C.c_int32, [c_ds_p]),
    "ppca_b200_llks": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, c_dp]),
    "ppca_b200_llk": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, c_dp]),
    "ppca_b200_infer": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, c_dp, c_dp]),
    "ppca_b200_covariance_full": (C.c_int32, [c_ctx_p, C.c_int64, C.c_int32, C.c_int32, c_dp, C.c_double, c_dp, c_ds_p, c_dp]),
    "ppca_b200_covariance_diagonal": (
        C.c_int32,
        [c_ctx_p, C.c_int64, C.c_int32, C.c_int32, c_dp, C.c_double, c_dp, c_ds_p, C.POINTER(c_ds_p)],
    ),
    "ppca_b200_smooth": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(c_ds_p)]),
    "ppca_b200_extrapolate": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(c_ds_p)]),
    "ppca_b200_reconstruct": (
        C.c_int32,
        [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, C.c_int32, c_ds_p, c_dp, C.POINTER(c_ds_p)],
    ),
    "ppca_b200_iterate": (
        C.c_int32,
        [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(CPrior), c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_iterate_host": (
        C.c_int32,
        [c_ctx_p, c_dp, C.c_int64, C.c_int32, c_dp, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(CPrior), c_dp, c_dp,
         c_dp, c_dp],
    ),
    "ppca_b200_pack_host": (
        C.c_int32, [c_dp, C.c_int64, C.c_int32, c_dp, C.POINTER(C.c_int64), C.POINTER(C.c_uint32)]),
    "ppca_b200_iterate_packed_host": (
        C.c_int32,
        [c_ctx_p, c_dp, C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.c_int64, C.c_int32, c_dp, C.c_int32, c_dp, c_dp,
         C.c_double, C.POINTER(CPrior), c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_iterate_packed_host_sharded": (
        C.c_int32,
        [c_ctx_p, c_dp, C.POINTER(C.c_int64), C.POINTER(C.c_uint32), C.c_int64, C.c_int32, c_dp, C.c_int32, c_dp, c_dp,
         C.c_double, C.POINTER(CPrior), c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_em_stats_host": (
        C.c_int32, [c_ctx_p, c_dp, C.c_int64, C.c_int32, c_dp, C.c_int32, c_dp, c_dp, C.c_double, C.c_void_p]),
    "ppca_b200_reconstruct_host": (
        C.c_int32, [c_ctx_p, c_dp, C.c_int64, C.c_int32, C.c_int32, c_dp, c_dp, C.c_double, C.c_int32, c_dp, c_dp]),
    "ppca_b200_host_register": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "ppca_b200_host_unregister": (C.c_int32, [C.c_void_p]),
    "ppca_b200_em_stats_len": (C.c_int64, [C.c_int32, C.c_int32]),
    "ppca_b200_em_stats": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, C.c_void_p]),
    "ppca_b200_em_finish": (
        C.c_int32,
        [c_ctx_p, C.c_int32, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(CPrior), C.c_void_p, c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_mix_llks": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "ppca_b200_mix_llk": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "ppca_b200_mix_infer_cluster": (C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "ppca_b200_mix_smooth": (
        C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.POINTER(c_ds_p)]),
    "ppca_b200_mix_extrapolate": (
        C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.POINTER(c_ds_p)]),
    "ppca_b200_mix_iterate": (
        C.c_int32,
        [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.POINTER(CPrior), c_dp, c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_mix_posteriors": (
        C.c_int32, [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.c_void_p, c_dp, c_dp]),
    "ppca_b200_comm_unique_id": (C.c_int32, [C.POINTER(C.c_uint8)]),
    "ppca_b200_comm_init": (C.c_int32, [c_ctx_p, C.POINTER(C.c_uint8), C.c_int32, C.c_int32]),
    "ppca_b200_comm_destroy": (C.c_int32, [c_ctx_p]),
    "ppca_b200_comm_allreduce": (C.c_int32, [c_ctx_p, C.c_void_p, C.c_int64, C.c_int32]),
    "ppca_b200_iterate_sharded": (
        C.c_int32,
        [c_ctx_p, c_ds_p, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(CPrior), c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_iterate_generated": (
        C.c_int32,
        [c_ctx_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_int32, c_dp, c_dp,
         C.c_double, C.POINTER(CPrior), C.c_int32, c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_iterate_host_sharded": (
        C.c_int32,
        [c_ctx_p, c_dp, C.c_int64, C.c_int32, c_dp, C.c_int32, c_dp, c_dp, C.c_double, C.POINTER(CPrior), c_dp, c_dp,
         c_dp, c_dp],
    ),
    "ppca_b200_mix_iterate_sharded": (
        C.c_int32,
        [c_ctx_p, c_ds_p, C.c_int32, c_ip, c_dp, c_dp, c_dp, c_dp, C.POINTER(CPrior), c_dp, c_dp, c_dp, c_dp, c_dp],
    ),
    "ppca_b200_mix_em_stats": (
        C.c_int32,
        [c_ctx_p, c_ds_p, C.c_int32, C.c_int32, C.c_int32, c_dp, c_dp, C.c_double, C.c_void_p, C.c_double, C.c_void_p],
    ),
}

_lib = None
_lib_lock = threading.Lock()


def lib():
    """Loads the CUDA library; raises (never falls back) if it is missing."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(nvcc, sm_100a). There is no CPU fallback."
                )
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in _PROTOTYPES.items():
                fn = getattr(handle, name)
                fn.restype = res
                fn.argtypes = args
            _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ppca_b200_last_error()
        raise NativeError(rc, msg.decode() if msg else "unknown")


def dptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(c_dp)


def f64(a, copy: bool = False) -> np.ndarray:
    out = np.ascontiguousarray(a, dtype=np.float64)
    if copy and out is a:
        out = out.copy()
    return out


class Context:
    """One device + one stream (ppca_b200_ctx)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self.device = int(device)
        self._h = c_ctx_p()
        check(lib().ppca_b200_ctx_create(self.device, C.c_void_p(stream) if stream else None, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def synchronize(self) -> None:
        check(lib().ppca_b200_ctx_synchronize(self._h))

    def stream_handle(self) -> int:
        """The cudaStream_t the context launches on (to order foreign work, e.g. an NCCL all-reduce, with it)."""
        out = C.c_void_p()
        check(lib().ppca_b200_ctx_stream(self._h, C.byref(out)))
        return int(out.value or 0)

    # -- sample-sharded EM: the NCCL communicator lives inside the library (ppca_b200_comm_*) --------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        check(lib().ppca_b200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().ppca_b200_comm_init(self._h, buf, int(rank), int(world)))
        self.comm_rank, self.comm_world = int(rank), int(world)

    def comm_destroy(self) -> None:
        check(lib().ppca_b200_comm_destroy(self._h))

    def comm_allreduce(self, dev_ptr: int, count: int, op: str = "sum") -> None:
        check(lib().ppca_b200_comm_allreduce(self._h, C.c_void_p(dev_ptr), int(count), {"sum": 0, "max": 1}[op]))

    def set_chunk(self, chunk: int) -> None:
        check(lib().ppca_b200_ctx_set_chunk(self._h, int(chunk)))

    def set_gemm(self, mode: str, slices: int = 6) -> None:
        """'dmma' (FP64 tensor cores), 'int8' (exact int8-sliced evaluation on mma.sync) or 'tc' (the same on
        tcgen05 with TMEM accumulators); see include/ppca_b200.h.  slices=4 with 'tc' is the FP32-class fast path."""
        check(lib().ppca_b200_ctx_set_gemm(self._h, {"dmma": 0, "int8": 1, "tc": 2}[mode], int(slices)))

    def set_guard(self, enabled: bool = True, eps_bits: int = 0) -> None:
        """Precision guard + ladder of the int8-sliced contractions (include/ppca_b200.h: ppca_b200_ctx_set_guard)."""
        check(lib().ppca_b200_ctx_set_guard(self._h, int(bool(enabled)), int(eps_bits)))

    def launch_count(self) -> int:
        out = C.c_int64(0)
        check(lib().ppca_b200_ctx_launch_count(self._h, C.byref(out)))
        return out.value

    VARIANTS = ("tc_smem_a", "tc_atm_feed", "tc_atm_drain", "tc_atm2", "imma", "dmma", "solve_reg8", "solve_reg16",
                "solve_reg32", "unused9", "solve_tile", "unused11", "solve_generic", "precision_retry",
                "tc_mix", "graph_replays")

    def variant_counts(self) -> Dict[str, int]:
        """Launches so far per shape-dependent kernel variant (ppca_b200_ctx_variant_counts): lets a test assert
        which contraction / solve kernel a case really ran."""
        out = (C.c_int64 * 16)()
        check(lib().ppca_b200_ctx_variant_counts(self._h, out))
        return dict(zip(self.VARIANTS, [int(v) for v in out]))

    def set_profiling(self, enabled: bool) -> None:
        check(lib().ppca_b200_ctx_set_profiling(self._h, int(bool(enabled))))

    def last_profile(self) -> Dict[str, float]:
        out = np.zeros(8)
        check(lib().ppca_b200_ctx_last_profile(self._h, dptr(out)))
        names = ["ksym", "gram", "proj", "solve", "moment", "cross_resid", "finish", "slice"]
        return dict(zip(names, out.tolist()))

    def close(self) -> None:
        if self._h:
            lib().ppca_b200_ctx_destroy(self._h)
            self._h = c_ctx_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


_contexts: Dict[int, Context] = {}
_ctx_lock = threading.Lock()


def default_device() -> int:
    for var in ("PPCA_B200_DEVICE", "LOCAL_RANK"):
        v = os.environ.get(var)
        if v is not None and v != "":
            return int(v)
    return 0


def get_context(device: Optional[int] = None) -> Context:
    """Per-process default context of a device (created on first use)."""
    dev = default_device() if device is None else int(device)
    with _ctx_lock:
        ctx = _contexts.get(dev)
        if ctx is None:
            ctx = Context(dev)
            _contexts[dev] = ctx
    return ctx


def set_context(ctx: Context) -> None:
    """Installs `ctx` as the default context of its device (e.g. one bound to torch's current stream)."""
    with _ctx_lock:
        _contexts[ctx.device] = ctx


def device_count() -> int:
    out = C.c_int32(0)
    check(lib().ppca_b200_device_count(C.byref(out)))
    return out.value
