"""Host-side mirror of the reference's Python extension module (`ppca_rs.ppca_rs`, src/python_bindings.rs).

Same class names, method names, argument meaning and error behaviour; the data-parallel work (E-step,
M-step statistics, log-likelihoods, reconstruction) runs in the CUDA engine through the C ABI
(include/ppca_b200.h).  Model parameters live on the host as numpy arrays (d k + d + 1 doubles), the
dataset lives on the device behind an opaque handle — exactly the split the reference has between Python
and its Rust side (Dataset is an opaque wrapper there too, src/python_bindings.rs:28-30).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterator, List, Optional, Sequence

import numpy as np

from . import _native as nat
from . import bincode


def _as_vector(x, what: str) -> np.ndarray:
    """src/utils.rs:10-23 to_nalgebra_vector: accepts 1 x n or n x 1 (1-D accepted as a convenience)."""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 1:
        return np.ascontiguousarray(a)
    if a.ndim == 2 and (a.shape[0] == 1 or a.shape[1] == 1):
        return np.ascontiguousarray(a.reshape(-1))
    raise ValueError(f"Expected column- or row- vector for {what}; got {'x'.join(map(str, a.shape))} matrix")


def _as_matrix(x, what: str) -> np.ndarray:
    a = np.asarray(x, dtype=np.float64)
    if a.ndim != 2:
        raise TypeError(f"{what} must be a 2-D float64 array")
    return np.ascontiguousarray(a)


# =================================================================================================
# Dataset  (src/python_bindings.rs:28-166 ; ppca/src/dataset.rs)
# =================================================================================================
def _device_view(obj, ndim: int, what: str):
    """(pointer, shape, strides in elements, keep-alive) of a float64 device array given through
    `__cuda_array_interface__` (torch, CuPy, Numba) or, failing that, `__dlpack__` (imported with torch)."""
    cai = getattr(obj, "__cuda_array_interface__", None)
    keep = obj
    if cai is None:
        if not hasattr(obj, "__dlpack__"):
            raise TypeError(f"{what} exposes neither __cuda_array_interface__ nor __dlpack__")
        import torch
        keep = torch.from_dlpack(obj)
        if not keep.is_cuda:
            raise ValueError(f"{what} is not on a CUDA device")
        cai = keep.__cuda_array_interface__
    if cai["typestr"] not in ("<f8", "=f8", "|f8"):
        raise TypeError(f"{what} must be float64 (got typestr {cai['typestr']})")
    shape = tuple(int(v) for v in cai["shape"])
    if len(shape) != ndim:
        raise ValueError(f"{what} must have {ndim} dimension(s), got shape {shape}")
    strides = cai.get("strides")
    if strides is None:
        strides, acc = [], 1
        for v in reversed(shape):
            strides.append(acc)
            acc *= max(v, 1)
        strides = tuple(reversed(strides))
    else:
        if any(int(v) % 8 for v in strides):
            raise ValueError(f"{what} has strides that are not multiples of 8 bytes")
        strides = tuple(int(v) // 8 for v in strides)
    ptr = int(cai["data"][0]) if cai["data"][0] is not None else 0
    return C.c_void_p(ptr), shape, strides, keep


def _producer_stream():
    """The CUDA stream the caller's framework is currently enqueueing on (torch's current stream when torch is loaded and
    CUDA is up; the legacy default stream otherwise): device ingestion is ordered after it."""
    import sys
    torch = sys.modules.get("torch")
    try:
        if torch is not None and torch.cuda.is_available() and torch.cuda.is_initialized():
            return C.c_void_p(int(torch.cuda.current_stream().cuda_stream))
    except Exception:
        pass
    return C.c_void_p(0)


class Dataset:
    """A dataset of samples with potentially missing values, resident on the GPU.

    `Dataset(ndarray, weights=None)`: `ndarray` is (n_samples, n_features) float64; every non-finite entry
    (NaN, +-inf) is a missing value (dataset.rs:19-22).
    """

    def __init__(self, ndarray, weights=None, *, _handle=None, _ctx=None):
        self._ctx = _ctx or nat.get_context()
        if _handle is not None:
            self._h = _handle
            return
        x = _as_matrix(ndarray, "ndarray")
        n, d = x.shape
        w = None
        if weights is not None:
            w = nat.f64(np.asarray(weights, dtype=np.float64).reshape(-1))
            if w.shape[0] != n:  # dataset.rs:163 assert_eq!(data.len(), weights.len())
                raise ValueError(f"weights has {w.shape[0]} entries for {n} samples")
        if d < 1:
            raise ValueError("dataset needs at least one output dimension")
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_dataset_from_host(self._ctx.handle, nat.dptr(x), n, d, nat.dptr(w), C.byref(h)))
        self._h = h

    # -- construction helpers --------------------------------------------------------------------
    @classmethod
    def _wrap(cls, handle, ctx) -> "Dataset":
        return cls(None, _handle=handle, _ctx=ctx)

    @classmethod
    def synthetic(cls, n: int, d: int, k_true: int, sigma_true: float = 0.1, mask_prob: float = 0.2,
                  n_components: int = 1, seed: int = 20240531, ctx: Optional[nat.Context] = None,
                  row_begin: int = 0) -> "Dataset":
        """Device-generated data with the reference sampler's semantics (ppca_model.rs:164-191).  `row_begin`: rows
        [row_begin, row_begin + n) of the dataset — the ranks of a sharded job pass one seed and their own row range."""
        ctx = ctx or nat.get_context()
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_dataset_synthetic_rows(ctx.handle, int(row_begin), int(n), int(d), int(k_true),
                                                             float(sigma_true), float(mask_prob), int(n_components),
                                                             int(seed), C.byref(h)))
        return cls._wrap(h, ctx)

    @classmethod
    def from_device(cls, array, weights=None, ctx: Optional[nat.Context] = None) -> "Dataset":
        """Zero-copy ingestion of a matrix that already lives on the GPU (ppca_b200_dataset_from_device): `array` (and
        `weights`) is anything exposing `__cuda_array_interface__` or `__dlpack__` — a torch.Tensor, a CuPy / Numba array —
        of dtype float64 and shape (n_samples, n_features), rows contiguous (any row stride).  Non-finite entries are the
        missing values, exactly as in `Dataset(ndarray)`; nothing crosses the PCIe bus.  The reference has no such path
        (it copies numpy -> Rust element by element, src/python_bindings.rs:41-54)."""
        ctx = ctx or nat.get_context()
        ptr, shape, strides, keep = _device_view(array, 2, "array")
        n, d = shape
        if strides[1] != 1 and n * d > 0:
            raise ValueError("array rows must be contiguous (unit stride along the feature axis)")
        if d < 1:
            raise ValueError("dataset needs at least one output dimension")
        row_stride = strides[0] if n > 1 else max(d, strides[0])
        if row_stride < d:
            raise ValueError("array rows overlap (row stride shorter than a row)")
        wptr = None
        if weights is not None:
            wptr, wshape, wstrides, wkeep = _device_view(weights, 1, "weights")
            if wshape[0] != n:  # dataset.rs:163
                raise ValueError(f"weights has {wshape[0]} entries for {n} samples")
            if wstrides[0] != 1 and n > 1:
                raise ValueError("weights must be contiguous")
            keep = (keep, wkeep)
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_dataset_from_device(ctx.handle, ptr, n, d, row_stride, wptr, _producer_stream(),
                                                          C.byref(h)))
        del keep
        return cls._wrap(h, ctx)

    def to_torch(self):
        """`numpy()` without leaving the GPU: a (n, d) float64 torch.Tensor on the dataset's device, NaN at the masked
        slots (ppca_b200_dataset_to_device)."""
        import torch
        n, d = len(self), self._output_size()
        out = torch.empty((n, d), dtype=torch.float64, device=f"cuda:{self._ctx.device}")
        if n:
            torch.cuda.current_stream(out.device).synchronize()  # the allocation may reuse memory with pending work
            nat.check(nat.lib().ppca_b200_dataset_to_device(self._ctx.handle, self._h, 0, n, out.data_ptr()))
        return out

    def __dlpack__(self, stream=None):
        return self.to_torch().__dlpack__(stream=stream)

    def __dlpack_device__(self):
        return (2, self._ctx.device)  # kDLCUDA

    def __del__(self):  # pragma: no cover
        try:
            if getattr(self, "_h", None):
                nat.lib().ppca_b200_dataset_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- reference API ---------------------------------------------------------------------------
    @staticmethod
    def load(data: bytes) -> "Dataset":
        try:
            x, w = bincode.load_dataset(bytes(data))
        except ValueError as e:
            raise Exception(str(e))
        if x.shape[0] == 0:
            raise Exception("cannot load an empty dataset (output size unknown)")
        return Dataset(x, w)

    def dump(self) -> bytes:
        return bincode.dump_dataset(self.numpy(), self.weights())

    def numpy(self) -> np.ndarray:
        n, d = len(self), self._output_size()
        out = np.empty((n, d), dtype=np.float64)
        nat.check(nat.lib().ppca_b200_dataset_to_host(self._ctx.handle, self._h, 0, n, nat.dptr(out)))
        return out

    def __len__(self) -> int:
        out = C.c_int64(0)
        nat.check(nat.lib().ppca_b200_dataset_len(self._h, C.byref(out)))
        return out.value

    def _output_size(self) -> int:
        out = C.c_int32(0)
        nat.check(nat.lib().ppca_b200_dataset_output_size(self._h, C.byref(out)))
        return out.value

    def output_size(self) -> Optional[int]:
        """dataset.rs:189-191: None for an empty dataset."""
        return self._output_size() if len(self) > 0 else None

    def empty_dimensions(self) -> List[int]:
        d = self._output_size()
        out = np.zeros(d, dtype=np.uint8)
        nat.check(nat.lib().ppca_b200_dataset_empty_dimensions(self._ctx.handle, self._h,
                                                               out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return [int(i) for i in np.nonzero(out)[0]]

    def weights(self) -> np.ndarray:
        out = np.empty(len(self), dtype=np.float64)
        nat.check(nat.lib().ppca_b200_dataset_weights(self._ctx.handle, self._h, nat.dptr(out)))
        return out

    def with_weights(self, weights) -> "Dataset":
        """dataset.rs:171-176: same samples (shared on the device), new weights."""
        w = nat.f64(np.asarray(weights, dtype=np.float64).reshape(-1))
        if w.shape[0] != len(self):
            raise ValueError("weights length does not match dataset length")
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_dataset_with_weights(self._ctx.handle, self._h, nat.dptr(w), C.byref(h)))
        return Dataset._wrap(h, self._ctx)

    def _slice(self, row0: int, nrows: int) -> "Dataset":
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_dataset_slice(self._ctx.handle, self._h, int(row0), int(nrows), C.byref(h)))
        return Dataset._wrap(h, self._ctx)

    def chunks(self, chunks: int) -> "DatasetChunks":
        return DatasetChunks(self, chunks)

    @staticmethod
    def concat(list: Sequence["Dataset"]) -> "Dataset":
        items = [d for d in list]
        if not items:
            raise ValueError("cannot concatenate an empty list of datasets (output size unknown)")
        ctx = items[0]._ctx
        arr = (nat.c_ds_p * len(items))(*[d._h for d in items])
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_dataset_concat(ctx.handle, arr, len(items), C.byref(h)))
        return Dataset._wrap(h, ctx)

    # pickling (the reference pickles through dump/load)
    def __getstate__(self):
        return self.dump()

    def __setstate__(self, state):
        other = Dataset.load(state)
        self._ctx = other._ctx
        self._h = other._h
        other._h = None

    def __reduce__(self):
        return (_load_dataset, (self.dump(),))


def _load_dataset(data: bytes) -> Dataset:
    return Dataset.load(data)


class GeneratedDataset:
    """Out-of-core SYNTHETIC dataset: rows [row_begin, row_begin + n) of what `Dataset.synthetic(N, d, k_true, sigma_true,
    mask_prob, seed=seed)` would hold, never stored — every EM step regenerates each chunk on the device from the
    counter-based RNG and consumes it (`ppca_b200_iterate_generated`).  This is how BASELINE configs[2] (N = 100 M,
    d = 2048: 1.6 TB) runs at kernel speed on 1 ... 8 GPUs; a rank of a sharded job passes its own row range.
    Accepted by `PPCAModel.iterate / iterate_with_prior / _iterate`."""

    def __init__(self, n: int, d: int, k_true: int, sigma_true: float = 0.1, mask_prob: float = 0.2,
                 seed: int = 20240531, row_begin: int = 0, ctx: Optional[nat.Context] = None):
        if n < 0 or d < 1 or k_true < 1 or row_begin < 0:
            raise ValueError("bad synthetic shape")
        self._ctx = ctx or nat.get_context()
        self.n, self.d, self.k_true = int(n), int(d), int(k_true)
        self.sigma_true, self.mask_prob, self.seed, self.row_begin = float(sigma_true), float(mask_prob), int(seed), int(row_begin)

    def __len__(self) -> int:
        return self.n

    def _output_size(self) -> int:
        return self.d

    def output_size(self) -> int:
        return self.d

    def materialize(self) -> "Dataset":
        """The same rows as a resident Dataset (for tests / small sizes)."""
        return Dataset.synthetic(self.n, self.d, self.k_true, self.sigma_true, self.mask_prob, 1, self.seed, self._ctx,
                                 row_begin=self.row_begin)


class HostDataset:
    """Out-of-core dataset: the samples stay in (page-locked) HOST memory and are streamed through the GPU block by
    block on every EM step (`ppca_b200_iterate_host`), the H2D copy of one block overlapping the kernels of the
    previous one.  For data that does not fit the device (BASELINE config 3 is 1.64 TB) or is visited once.
    Accepted by `PPCAModel.iterate / iterate_with_prior` and `PPCATrainer`; same numbers as `Dataset`."""

    def __init__(self, ndarray, weights=None, *, pin: bool = True, ctx: Optional[nat.Context] = None,
                 packed: bool = False):
        """packed=True keeps a compact copy of the samples in page-locked memory - only the OBSERVED values, row
        offsets and the mask words (ppca_b200_pack_host) - and streams that instead of the full matrix: the bytes that
        cross PCIe per EM step drop by the missing fraction (same results, bit for bit)."""
        self._ctx = ctx or nat.get_context()
        self._x = _as_matrix(ndarray, "ndarray")
        n, d = self._x.shape
        if d < 1:
            raise ValueError("dataset needs at least one output dimension")
        self._w = None
        if weights is not None:
            self._w = nat.f64(np.asarray(weights, dtype=np.float64).reshape(-1))
            if self._w.shape[0] != n:  # dataset.rs:163
                raise ValueError(f"weights has {self._w.shape[0]} entries for {n} samples")
        self._pinned = []
        self._packed = None
        if packed and n > 0:
            dw = (d + 31) // 32
            rowptr = np.empty(n + 1, dtype=np.int64)
            lib = nat.lib()
            nat.check(lib.ppca_b200_pack_host(nat.dptr(self._x), n, d, None, rowptr.ctypes.data_as(C.POINTER(C.c_int64)), None))
            vals = np.empty(max(int(rowptr[n]), 1))
            maskw = np.empty((n, dw), dtype=np.uint32)
            nat.check(lib.ppca_b200_pack_host(nat.dptr(self._x), n, d, nat.dptr(vals),
                                              rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
                                              maskw.ctypes.data_as(C.POINTER(C.c_uint32))))
            self._packed = (vals, rowptr, maskw)
        if pin and n > 0:
            to_pin = (self._w,) + self._packed if self._packed is not None else (self._x, self._w)
            for a in to_pin:
                if a is not None and a.nbytes > 0:
                    # page-locking can be refused (RLIMIT_MEMLOCK, already-registered ranges): the data then streams
                    # from pageable memory - same results, lower PCIe rate
                    if nat.lib().ppca_b200_host_register(C.c_void_p(a.ctypes.data), a.nbytes) == 0:
                        self._pinned.append(a.ctypes.data)

    def __del__(self):  # pragma: no cover
        try:
            for ptr in getattr(self, "_pinned", []):
                nat.lib().ppca_b200_host_unregister(C.c_void_p(ptr))
            self._pinned = []
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._x.shape[0])

    def _output_size(self) -> int:
        return int(self._x.shape[1])

    def output_size(self) -> Optional[int]:
        return self._output_size() if len(self) > 0 else None

    def numpy(self) -> np.ndarray:
        out = self._x.copy()
        out[~np.isfinite(out)] = np.nan
        return out

    def weights(self) -> np.ndarray:
        return np.ones(len(self)) if self._w is None else self._w.copy()

    def empty_dimensions(self) -> List[int]:
        """dataset.rs:194-222 (host side: the data is here)."""
        if len(self) == 0:
            return []
        return [int(i) for i in np.nonzero(~np.isfinite(self._x).any(axis=0))[0]]

    def resident(self) -> Dataset:
        """Uploads once and returns the device-resident Dataset."""
        return Dataset(self._x, self._w)


class DatasetChunks:
    """src/python_bindings.rs:136-166: stride = ceil(len / chunks), yields copies in order."""

    def __init__(self, dataset: Dataset, chunks: int):
        self.length = len(dataset)
        self.stride = int(math.ceil(self.length / float(chunks))) if chunks > 0 else 0
        self.position = 0
        self.dataset = dataset

    def __iter__(self) -> Iterator[Dataset]:
        return self

    def __next__(self) -> Dataset:
        if self.position < self.length and self.stride > 0:
            end = min(self.length, self.position + self.stride)
            out = self.dataset._slice(self.position, end - self.position)
            self.position += self.stride
            return out
        raise StopIteration


# =================================================================================================
# Prior  (src/python_bindings.rs:168-201 ; ppca/src/prior.rs)
# =================================================================================================
class Prior:
    """A prior for the PPCA model (mean N(mu0, S0), inverse-gamma noise, ridge on the transform rows)."""

    def __init__(self):
        self.mean: Optional[np.ndarray] = None
        self.mean_covariance: Optional[np.ndarray] = None
        self.mean_precision: Optional[np.ndarray] = None
        self.isotropic_noise_alpha: Optional[float] = None
        self.isotropic_noise_beta: Optional[float] = None
        self.transformation_precision: float = 0.0

    def _clone(self) -> "Prior":
        p = Prior()
        p.__dict__.update(self.__dict__)
        return p

    def with_mean_prior(self, mean, mean_covariance) -> "Prior":
        m = _as_vector(mean, "mean")
        cov = _as_matrix(mean_covariance, "mean_covariance")
        if cov.shape != (m.shape[0], m.shape[0]):  # prior.rs:33-34
            raise ValueError("mean covariance shape does not match mean length")
        try:
            prec = np.linalg.inv(cov)  # prior.rs:36-41 try_inverse().expect(...)
        except np.linalg.LinAlgError:
            raise ValueError("mean covariance should be invertible")
        p = self._clone()
        p.mean, p.mean_covariance, p.mean_precision = m, cov, np.ascontiguousarray(prec)
        return p

    def with_isotropic_noise_prior(self, alpha: float, beta: float) -> "Prior":
        if not (alpha >= 0.0 and beta >= 0.0):  # prior.rs:50-51
            raise ValueError("alpha and beta must be non-negative")
        p = self._clone()
        p.isotropic_noise_alpha, p.isotropic_noise_beta = float(alpha), float(beta)
        return p

    def with_transformation_precision(self, precision: float) -> "Prior":
        if not precision >= 0.0:  # prior.rs:61
            raise ValueError("precision must be non-negative")
        p = self._clone()
        p.transformation_precision = float(precision)
        return p

    def _c(self, d: int):
        """Returns (CPrior, keepalive)."""
        pr = nat.CPrior()
        keep = []
        pr.has_mean_prior = 0
        if self.mean is not None:
            if self.mean.shape[0] != d:
                raise ValueError("mean prior length does not match the model output size")
            keep = [self.mean, self.mean_precision]
            pr.has_mean_prior = 1
            pr.mean = nat.dptr(self.mean)
            pr.mean_precision = nat.dptr(self.mean_precision)
        pr.has_isotropic_noise_prior = int(self.isotropic_noise_alpha is not None)
        pr.isotropic_noise_alpha = float(self.isotropic_noise_alpha or 0.0)
        pr.isotropic_noise_beta = float(self.isotropic_noise_beta or 0.0)
        pr.transformation_precision = float(self.transformation_precision)
        return pr, keep


# =================================================================================================
# PPCAModel  (src/python_bindings.rs:367-533 ; ppca/src/ppca_model.rs)
# =================================================================================================
class PPCAModel:
    """x ~ N(0, I_k);  y = C x + mu + noise;  noise ~ N(0, sigma^2 I_d)   (ppca_model.rs:24-38)."""

    def __init__(self, isotropic_noise: float, transform, mean):
        self._sigma = float(isotropic_noise)
        self._C = _as_matrix(transform, "transform")
        self._mu = _as_vector(mean, "mean")
        if self._mu.shape[0] != self._C.shape[0]:
            raise ValueError("mean length does not match the number of rows of transform")

    # -- getters ---------------------------------------------------------------------------------
    @property
    def output_size(self) -> int:
        return int(self._C.shape[0])

    @property
    def state_size(self) -> int:
        return int(self._C.shape[1])

    @property
    def n_parameters(self) -> int:
        return 1 + self.state_size * self.output_size + self._mu.shape[0]  # ppca_model.rs:107-109

    @property
    def singular_values(self) -> np.ndarray:
        return np.sqrt(np.linalg.norm(self._C, axis=0))  # ppca_model.rs:113-121: sqrt of the column NORM

    @property
    def transform(self) -> np.ndarray:
        return self._C.copy()

    @property
    def isotropic_noise(self) -> float:
        return self._sigma

    @property
    def mean(self) -> np.ndarray:
        return self._mu.copy()

    # -- construction ----------------------------------------------------------------------------
    @staticmethod
    def init(state_size: int, dataset: Dataset, seed: Optional[int] = None) -> "PPCAModel":
        """ppca_model.rs:51-70: sigma = 1, mu = 0, C ~ N(0, 1) with the rows of empty dimensions zeroed."""
        if len(dataset) == 0:
            raise ValueError("dataset must not be empty")
        d = dataset.output_size()
        rng = np.random.default_rng(seed)
        C0 = rng.standard_normal((d, int(state_size)))
        for i in dataset.empty_dimensions():
            C0[i, :] = 0.0
        return PPCAModel(1.0, C0, np.zeros(d))

    @staticmethod
    def load(data: bytes) -> "PPCAModel":
        try:
            sigma, transform, mean = bincode.load_model(bytes(data))
        except ValueError as e:
            raise Exception(str(e))
        return PPCAModel(sigma, transform, mean)

    def dump(self) -> bytes:
        return bincode.dump_model(self._sigma, self._C, self._mu)

    def __repr__(self) -> str:
        return (f"PPCAModel(isotropic_noise={self._sigma}, transform=array({self._C}, dtype=\"float32\"), "
                f"mean=narray({self._mu}, dtype=\"float32\"))")

    # -- hot path --------------------------------------------------------------------------------
    def _check(self, dataset: Dataset) -> None:
        if len(dataset) > 0 and dataset._output_size() != self.output_size:  # output_covariance.rs:124
            raise ValueError(f"dataset output size {dataset._output_size()} != model output size {self.output_size}")
        if self.state_size < 1:
            raise ValueError("state_size 0 is not supported by the B200 engine")

    def llk(self, dataset: Dataset) -> float:
        self._check(dataset)
        if isinstance(dataset, HostDataset):
            return float(np.dot(self._host_pass(dataset, None, True)[1], dataset.weights())) if len(dataset) else 0.0
        out = C.c_double(0.0)
        nat.check(nat.lib().ppca_b200_llk(dataset._ctx.handle, dataset._h, self.state_size, nat.dptr(self._C),
                                          nat.dptr(self._mu), self._sigma, C.byref(out)))
        return out.value

    def _host_pass(self, dataset: "HostDataset", extrapolate: Optional[bool], want_llks: bool):
        """Streams a HostDataset through ppca_b200_reconstruct_host; returns (HostDataset | None, llks | None)."""
        n, d = len(dataset), self.output_size
        out = None
        if extrapolate is not None:
            out = HostDataset(np.empty((n, d)), None if dataset._w is None else dataset._w.copy(),
                              pin=bool(dataset._pinned), ctx=dataset._ctx)
        llks = np.empty(n) if want_llks else None
        nat.check(nat.lib().ppca_b200_reconstruct_host(dataset._ctx.handle, nat.dptr(dataset._x), n, d, self.state_size,
                                                       nat.dptr(self._C), nat.dptr(self._mu), self._sigma,
                                                       int(bool(extrapolate)), nat.dptr(out._x) if out else None,
                                                       nat.dptr(llks)))
        return out, llks

    def llks(self, dataset: Dataset) -> np.ndarray:
        self._check(dataset)
        if isinstance(dataset, HostDataset):
            return self._host_pass(dataset, None, True)[1]
        out = np.empty(len(dataset))
        nat.check(nat.lib().ppca_b200_llks(dataset._ctx.handle, dataset._h, self.state_size, nat.dptr(self._C),
                                           nat.dptr(self._mu), self._sigma, nat.dptr(out)))
        return out

    def sample(self, dataset_size: int, mask_prob: float, seed: Optional[int] = None) -> Dataset:
        """ppca_model.rs:164-191, generated on the device (counter-based RNG).  Unseeded like the reference unless
        `seed` is given (extension)."""
        if not 0.0 <= mask_prob <= 1.0:
            raise ValueError("invalid mask probability")
        if self.state_size < 1:
            raise ValueError("state_size 0 is not supported by the B200 engine")
        if seed is None:
            seed = int(np.random.default_rng().integers(0, 2 ** 63))
        ctx = nat.get_context()
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_model_sample(ctx.handle, int(dataset_size), self.output_size, self.state_size,
                                                   nat.dptr(self._C), nat.dptr(self._mu), self._sigma, float(mask_prob),
                                                   int(seed), C.byref(h)))
        return Dataset._wrap(h, ctx)

    def infer(self, dataset: Dataset) -> "InferredMasked":
        self._check(dataset)
        n, k = len(dataset), self.state_size
        states = np.empty((n, k))
        covs = np.empty((n, k, k))
        nat.check(nat.lib().ppca_b200_infer(dataset._ctx.handle, dataset._h, k, nat.dptr(self._C), nat.dptr(self._mu),
                                            self._sigma, nat.dptr(states), nat.dptr(covs)))
        return InferredMasked(states, covs, self)

    def _recon(self, dataset: Dataset, fn) -> Dataset:
        self._check(dataset)
        if isinstance(dataset, HostDataset):  # host in, host out, streamed
            return self._host_pass(dataset, fn is nat.lib().ppca_b200_extrapolate, False)[0]
        h = nat.c_ds_p()
        nat.check(fn(dataset._ctx.handle, dataset._h, self.state_size, nat.dptr(self._C), nat.dptr(self._mu),
                     self._sigma, C.byref(h)))
        return Dataset._wrap(h, dataset._ctx)

    def reconstruct(self, dataset: Dataset, extrapolate: bool = True, out: Optional[Dataset] = None,
                    with_llks: bool = False):
        """smooth / extrapolate and (with_llks) the per-sample log-likelihoods from ONE E-step (ppca_b200_reconstruct).
        `out` = a dataset an earlier smooth / extrapolate / reconstruct call returned for an input of the same shape: it
        is overwritten in place and returned (no allocation).  Returns the dataset, or (dataset, llks).  Extension of the
        reference API (it calls extrapolate and llks separately, ppca_model.rs:152-159,254-261)."""
        self._check(dataset)
        if isinstance(dataset, HostDataset):
            res, llks = self._host_pass(dataset, extrapolate, with_llks)
            return (res, llks) if with_llks else res
        llks = np.empty(len(dataset)) if with_llks else None
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_reconstruct(dataset._ctx.handle, dataset._h, self.state_size, nat.dptr(self._C),
                                                  nat.dptr(self._mu), self._sigma, int(bool(extrapolate)),
                                                  out._h if out is not None else None, nat.dptr(llks), C.byref(h)))
        res = out if out is not None else Dataset._wrap(h, dataset._ctx)
        return (res, llks) if with_llks else res

    def smooth(self, dataset: Dataset) -> Dataset:
        return self._recon(dataset, nat.lib().ppca_b200_smooth)

    # readme.md:62 calls `smooth` filter_extrapolate
    filter_extrapolate = smooth

    def extrapolate(self, dataset: Dataset) -> Dataset:
        return self._recon(dataset, nat.lib().ppca_b200_extrapolate)

    def _iterate(self, dataset: Dataset, prior: Optional[Prior], sharded: bool = False):
        """Returns (new model, llk of THIS model on dataset) — the E-step yields the latter for free.
        sharded=True: `dataset` is this rank's shard and the context carries a communicator (Context.comm_init): the
        statistics are all-reduced inside the library (ppca_b200_iterate_sharded / _host_sharded), every rank returns
        the same model and the GLOBAL log-likelihood."""
        self._check(dataset)
        d, k = self.output_size, self.state_size
        C_out = np.empty((d, k))
        mu_out = np.empty(d)
        s_out = C.c_double(0.0)
        llk = C.c_double(0.0)
        pr_ref, keep = None, None
        if prior is not None:
            pr, keep = prior._c(d)
            pr_ref = C.byref(pr)
        lib = nat.lib()
        if isinstance(dataset, GeneratedDataset):
            nat.check(lib.ppca_b200_iterate_generated(
                dataset._ctx.handle, dataset.row_begin, dataset.n, d, dataset.k_true, dataset.sigma_true, dataset.mask_prob,
                dataset.seed, k, nat.dptr(self._C), nat.dptr(self._mu), self._sigma, pr_ref, 1 if sharded else 0,
                nat.dptr(C_out), nat.dptr(mu_out), C.byref(s_out), C.byref(llk)))
            return PPCAModel(s_out.value, C_out, mu_out), llk.value
        if isinstance(dataset, HostDataset) and dataset._packed is not None:
            vals, rowptr, maskw = dataset._packed
            fn = lib.ppca_b200_iterate_packed_host_sharded if sharded else lib.ppca_b200_iterate_packed_host
            nat.check(fn(dataset._ctx.handle, nat.dptr(vals), rowptr.ctypes.data_as(C.POINTER(C.c_int64)),
                         maskw.ctypes.data_as(C.POINTER(C.c_uint32)), len(dataset), d, nat.dptr(dataset._w), k,
                         nat.dptr(self._C), nat.dptr(self._mu), self._sigma, pr_ref, nat.dptr(C_out), nat.dptr(mu_out),
                         C.byref(s_out), C.byref(llk)))
            return PPCAModel(s_out.value, C_out, mu_out), llk.value
        if isinstance(dataset, HostDataset):
            fn = lib.ppca_b200_iterate_host_sharded if sharded else lib.ppca_b200_iterate_host
            nat.check(fn(dataset._ctx.handle, nat.dptr(dataset._x), len(dataset), d, nat.dptr(dataset._w), k,
                         nat.dptr(self._C), nat.dptr(self._mu), self._sigma, pr_ref, nat.dptr(C_out), nat.dptr(mu_out),
                         C.byref(s_out), C.byref(llk)))
            return PPCAModel(s_out.value, C_out, mu_out), llk.value
        fn = lib.ppca_b200_iterate_sharded if sharded else lib.ppca_b200_iterate
        nat.check(fn(dataset._ctx.handle, dataset._h, k, nat.dptr(self._C), nat.dptr(self._mu), self._sigma, pr_ref,
                     nat.dptr(C_out), nat.dptr(mu_out), C.byref(s_out), C.byref(llk)))
        return PPCAModel(s_out.value, C_out, mu_out), llk.value

    def iterate_with_prior(self, dataset: Dataset, prior: Prior) -> "PPCAModel":
        return self._iterate(dataset, prior)[0]

    def iterate(self, dataset: Dataset) -> "PPCAModel":
        return self._iterate(dataset, None)[0]

    def to_canonical(self) -> "PPCAModel":
        """ppca_model.rs:398-425: C <- U S (singular values descending), columns times signum(sum(column))."""
        if self.state_size == 0:
            return PPCAModel(self._sigma, self._C, self._mu)
        u, s, _ = np.linalg.svd(self._C, full_matrices=False)
        new_c = u * s
        sums = new_c.sum(axis=0)
        sign = np.where(np.signbit(sums), -1.0, 1.0)  # f64::signum: signum(+0.0) = 1, signum(-0.0) = -1
        sign = np.where(np.isnan(sums), np.nan, sign)
        return PPCAModel(self._sigma, new_c * sign, self._mu)

    # -- pickling (src/python_bindings.rs:513-532) -------------------------------------------------
    def __getstate__(self):
        return self.dump()

    def __setstate__(self, state):
        self._sigma, self._C, self._mu = bincode.load_model(bytes(state))

    def __getnewargs__(self):
        return (self._sigma, self._C.copy(), self._mu.copy())


# =================================================================================================
# InferredMasked  (src/python_bindings.rs:203-365 ; ppca_model.rs:428-626)
# Per-sample convenience algebra on the inferred posteriors.  Host numpy: SURVEY.md §8(f) rank 1 ("next").
# =================================================================================================
class InferredMasked:
    def __init__(self, states: np.ndarray, covariances: np.ndarray, model: Optional["PPCAModel"] = None):
        self._states = states
        self._covs = covariances
        self._model = model  # the reference struct keeps a clone of the model (ppca_model.rs:428-432)

    def __len__(self) -> int:
        return self._states.shape[0]

    def states(self) -> np.ndarray:
        if len(self) == 0:
            return np.zeros((0, 0))
        return self._states.copy()

    def covariances(self) -> List[np.ndarray]:
        return [c.copy() for c in self._covs]

    def smoothed(self, ppca: PPCAModel) -> Dataset:
        return Dataset(self._states @ ppca._C.T + ppca._mu)  # ppca_model.rs:454-456

    def extrapolated(self, ppca: PPCAModel, dataset: Dataset) -> Dataset:
        x = dataset.numpy()
        sm = self._states @ ppca._C.T + ppca._mu
        return Dataset(np.where(np.isfinite(x), x, sm))  # ppca_model.rs:460-463

    def _cov_full(self, ppca: PPCAModel, dataset: Optional[Dataset]) -> np.ndarray:
        """sigma^2 I + C Sigma_n C^T for every sample as one (n, d, d) array, on the device (ppca_b200_covariance_full);
        rows / columns of the dimensions `dataset` observed are zero when it is given (ppca_model.rs:471-477, 517-534)."""
        n, k, d = self._states.shape[0], ppca.state_size, ppca.output_size
        out = np.empty((n, d, d), dtype=np.float64)
        if n == 0:
            return out
        ctx = dataset._ctx if dataset is not None else nat.get_context()
        covs = nat.f64(self._covs)
        nat.check(nat.lib().ppca_b200_covariance_full(ctx.handle, n, d, k, nat.dptr(ppca._C), ppca._sigma, nat.dptr(covs),
                                                      dataset._h if dataset is not None else None, nat.dptr(out)))
        return out

    def smoothed_covariances(self, ppca: PPCAModel) -> List[np.ndarray]:
        return list(self._cov_full(ppca, None))  # ppca_model.rs:471-477

    def _cov_diag_dataset(self, ppca: PPCAModel, dataset: Optional[Dataset]) -> Dataset:
        """sigma^2 + c_i^T Sigma_n c_i for every (sample, dimension), on the device (ppca_b200_covariance_diagonal);
        slots `dataset` observed are 0 when it is given (ppca_model.rs:485-508, 542-577)."""
        n, k = self._states.shape[0], ppca.state_size
        ctx = dataset._ctx if dataset is not None else nat.get_context()
        covs = nat.f64(self._covs)
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_covariance_diagonal(ctx.handle, n, ppca.output_size, k, nat.dptr(ppca._C),
                                                          ppca._sigma, nat.dptr(covs),
                                                          dataset._h if dataset is not None else None, C.byref(h)))
        return Dataset._wrap(h, ctx)

    def _smoothed_cov_diag(self, ppca: PPCAModel) -> np.ndarray:
        return self._cov_diag_dataset(ppca, None).numpy()

    def smoothed_covariances_diagonal(self, ppca: PPCAModel) -> Dataset:
        return self._cov_diag_dataset(ppca, None)

    def extrapolated_covariances(self, ppca: PPCAModel, dataset: Dataset) -> List[np.ndarray]:
        if len(dataset) != len(self):
            raise ValueError("dataset length does not match the inferred batch")
        return list(self._cov_full(ppca, dataset))  # ppca_model.rs:517-534

    def extrapolated_covariances_diagonal(self, ppca: PPCAModel, dataset: Dataset) -> Dataset:
        return self._cov_diag_dataset(ppca, dataset)  # ppca_model.rs:542-577

    def posterior_sampler(self) -> "PosteriorSampler":
        return PosteriorSampler(self._states, self._covs, self._model)  # ppca_model.rs:581-592


def _fresh_seed(seed: Optional[int]) -> int:
    return int(np.random.default_rng().integers(0, 2 ** 63)) if seed is None else int(seed)


def _posterior_sample(models: Sequence["PPCAModel"], posteriors: Optional[np.ndarray], states: Sequence[np.ndarray],
                      covs: Sequence[np.ndarray], seed: Optional[int]) -> Dataset:
    """One posterior draw per inferred sample on the device (ppca_b200_posterior_sample)."""
    ctx = nat.get_context()
    m = len(models)
    n = int(states[0].shape[0])
    d = models[0].output_size
    ks = np.array([mm.state_size for mm in models], dtype=np.int32)
    Cs = nat.f64(np.concatenate([mm._C.reshape(-1) for mm in models]))
    mus = nat.f64(np.concatenate([mm._mu.reshape(-1) for mm in models]))
    sig = nat.f64(np.array([mm._sigma for mm in models], dtype=np.float64))
    st = [nat.f64(np.asarray(v, dtype=np.float64).reshape(n, -1)) for v in states]
    cv = [nat.f64(np.asarray(v, dtype=np.float64).reshape(n, -1)) for v in covs]
    sp = (nat.c_dp * m)(*[nat.dptr(v) for v in st])
    cp = (nat.c_dp * m)(*[nat.dptr(v) for v in cv])
    post = nat.f64(np.asarray(posteriors, dtype=np.float64).reshape(n, m)) if m > 1 else None
    h = nat.c_ds_p()
    nat.check(nat.lib().ppca_b200_posterior_sample(ctx.handle, n, d, m, ks.ctypes.data_as(nat.c_ip), nat.dptr(Cs),
                                                   nat.dptr(mus), nat.dptr(sig), nat.dptr(post), sp, cp,
                                                   _fresh_seed(seed), C.byref(h)))
    return Dataset._wrap(h, ctx)


class PosteriorSampler:
    """ppca_model.rs:595-626: x = noise + mean + C (state + L standard), L L^T = covariance.  Drawn on the device; the
    Cholesky factors are formed there too (a covariance that is not positive definite raises, as the reference's
    `expect("Cholesky decomposition failed")`, ppca_model.rs:582-586)."""

    def __init__(self, states: np.ndarray, covs: np.ndarray, model: Optional["PPCAModel"] = None):
        self._states, self._covs, self._model = states, covs, model

    def sample(self, model: Optional["PPCAModel"] = None, seed: Optional[int] = None) -> Dataset:
        """No arguments, as in the reference (src/python_bindings.rs:335-365); `seed` is an extension (the reference
        draws from thread_rng)."""
        model = model or self._model
        if model is None:
            raise ValueError("a PPCAModel is needed to sample outputs")
        return _posterior_sample([model], None, [self._states], [self._covs], seed)


# =================================================================================================
# PPCAMix  (src/python_bindings.rs:535-711 ; ppca/src/mix.rs)
# =================================================================================================
def _log_softmax(v: np.ndarray) -> np.ndarray:
    """mix.rs:14-18 robust_log_softmax."""
    mx = np.max(v)
    return v - mx - np.log(np.sum(np.exp(v - mx)))


class PPCAMix:
    def __init__(self, models: Sequence[PPCAModel], log_weights):
        models = list(models)
        lw = np.asarray(log_weights, dtype=np.float64).reshape(-1)
        if len(models) == 0:  # mix.rs:51
            raise ValueError("a PPCA mixture needs at least one model")
        if len(models) != lw.shape[0]:  # mix.rs:52
            raise ValueError("models and log_weights have different lengths")
        sizes = [m.output_size for m in models]
        if len(set(sizes)) != 1:  # mix.rs:58-64
            raise ValueError(f"Model output sizes are not the same: {sizes}")
        self._models = models
        self._logw = _log_softmax(lw)  # mix.rs:69

    @staticmethod
    def init(n_models: int, state_size: int, dataset: Dataset) -> "PPCAMix":
        return PPCAMix([PPCAModel.init(state_size, dataset) for _ in range(n_models)], np.zeros(n_models))

    @staticmethod
    def load(data: bytes) -> "PPCAMix":
        try:
            _, models, lw = bincode.load_mix(bytes(data))
        except ValueError as e:
            raise Exception(str(e))
        mix = PPCAMix.__new__(PPCAMix)
        mix._models = [PPCAModel(s, c, m) for s, c, m in models]
        mix._logw = np.asarray(lw, dtype=np.float64)  # stored already normalised
        return mix

    def dump(self) -> bytes:
        return bincode.dump_mix(self.output_size, [(m._sigma, m._C, m._mu) for m in self._models], self._logw)

    @property
    def output_size(self) -> int:
        return self._models[0].output_size

    @property
    def state_sizes(self) -> List[int]:
        return [m.state_size for m in self._models]

    @property
    def n_parameters(self) -> int:
        return sum(m.n_parameters for m in self._models) + len(self._models) - 1  # mix.rs:96-104

    @property
    def models(self) -> List[PPCAModel]:
        return list(self._models)

    @property
    def log_weights(self) -> np.ndarray:
        return self._logw.copy()

    @property
    def weights(self) -> np.ndarray:
        return np.exp(self._logw)

    def __repr__(self) -> str:
        return f"PPCAMix(models={self._models!r}, log_weights={self._logw!r})"

    # -- packing for the C ABI -------------------------------------------------------------------
    def _pack(self):
        ks = np.array(self.state_sizes, dtype=np.int32)
        if (ks < 1).any():
            raise ValueError("state_size 0 is not supported by the B200 engine")
        Cs = np.ascontiguousarray(np.concatenate([m._C.reshape(-1) for m in self._models]))
        mus = np.ascontiguousarray(np.stack([m._mu for m in self._models]))
        sig = np.array([m._sigma for m in self._models], dtype=np.float64)
        return ks, Cs, mus, sig, np.ascontiguousarray(self._logw)

    def _check(self, dataset: Dataset) -> None:
        if len(dataset) > 0 and dataset._output_size() != self.output_size:
            raise ValueError("dataset output size does not match the mixture output size")

    def _call(self, fn, dataset: Dataset, *tail):
        self._check(dataset)
        ks, Cs, mus, sig, lw = self._pack()
        nat.check(fn(dataset._ctx.handle, dataset._h, len(self._models), ks.ctypes.data_as(nat.c_ip), nat.dptr(Cs),
                     nat.dptr(mus), nat.dptr(sig), nat.dptr(lw), *tail))

    def llks(self, dataset: Dataset) -> np.ndarray:
        out = np.empty(len(dataset))
        self._call(nat.lib().ppca_b200_mix_llks, dataset, nat.dptr(out))
        return out

    def llk(self, dataset: Dataset) -> float:
        out = C.c_double(0.0)
        self._call(nat.lib().ppca_b200_mix_llk, dataset, C.byref(out))
        return out.value

    def sample(self, dataset_size: int, mask_probability: float, seed: Optional[int] = None) -> Dataset:
        """mix.rs:124-134 on the device (ppca_b200_mix_sample): component from exp(log_weights), then that model's
        sample_one.  Unseeded like the reference unless `seed` is given (extension)."""
        if not 0.0 <= mask_probability <= 1.0:
            raise ValueError("invalid mask probability")
        ctx = nat.get_context()
        ks, Cs, mus, sig, lw = self._pack()
        h = nat.c_ds_p()
        nat.check(nat.lib().ppca_b200_mix_sample(ctx.handle, int(dataset_size), self.output_size, len(self._models),
                                                 ks.ctypes.data_as(nat.c_ip), nat.dptr(Cs), nat.dptr(mus), nat.dptr(sig),
                                                 nat.dptr(lw), float(mask_probability), _fresh_seed(seed), C.byref(h)))
        return Dataset._wrap(h, ctx)

    def infer_cluster(self, dataset: Dataset) -> np.ndarray:
        out = np.empty((len(dataset), len(self._models)))
        self._call(nat.lib().ppca_b200_mix_infer_cluster, dataset, nat.dptr(out))
        return out

    def infer(self, dataset: Dataset) -> "InferredMaskedMix":
        log_post = self.infer_cluster(dataset)
        return InferredMaskedMix(log_post, [m.infer(dataset) for m in self._models], self)

    def smooth(self, dataset: Dataset) -> Dataset:
        h = nat.c_ds_p()
        self._call(nat.lib().ppca_b200_mix_smooth, dataset, C.byref(h))
        return Dataset._wrap(h, dataset._ctx)

    filter_extrapolate = smooth

    def extrapolate(self, dataset: Dataset) -> Dataset:
        h = nat.c_ds_p()
        self._call(nat.lib().ppca_b200_mix_extrapolate, dataset, C.byref(h))
        return Dataset._wrap(h, dataset._ctx)

    def _iterate(self, dataset: Dataset, prior: Optional[Prior], sharded: bool = False):
        self._check(dataset)
        ks, Cs, mus, sig, lw = self._pack()
        Cs_o, mus_o, sig_o, lw_o = np.empty_like(Cs), np.empty_like(mus), np.empty_like(sig), np.empty_like(lw)
        llk = C.c_double(0.0)
        pr_ref, keep = None, None
        if prior is not None:
            pr, keep = prior._c(self.output_size)
            pr_ref = C.byref(pr)
        fn = nat.lib().ppca_b200_mix_iterate_sharded if sharded else nat.lib().ppca_b200_mix_iterate
        nat.check(fn(dataset._ctx.handle, dataset._h, len(self._models), ks.ctypes.data_as(nat.c_ip), nat.dptr(Cs),
                     nat.dptr(mus), nat.dptr(sig), nat.dptr(lw), pr_ref, nat.dptr(Cs_o), nat.dptr(mus_o),
                     nat.dptr(sig_o), nat.dptr(lw_o), C.byref(llk)))
        d = self.output_size
        models, off = [], 0
        for j, k in enumerate(ks):
            models.append(PPCAModel(float(sig_o[j]), Cs_o[off:off + d * k].reshape(d, k).copy(), mus_o[j].copy()))
            off += d * k
        mix = PPCAMix.__new__(PPCAMix)
        mix._models = models
        mix._logw = lw_o
        return mix, llk.value

    def iterate_with_prior(self, dataset: Dataset, prior: Prior) -> "PPCAMix":
        return self._iterate(dataset, prior)[0]

    def iterate(self, dataset: Dataset) -> "PPCAMix":
        return self._iterate(dataset, None)[0]

    def to_canonical(self) -> "PPCAMix":
        mix = PPCAMix.__new__(PPCAMix)
        mix._models = [m.to_canonical() for m in self._models]
        mix._logw = self._logw.copy()
        return mix

    def __getstate__(self):
        return self.dump()

    def __setstate__(self, state):
        other = PPCAMix.load(state)
        self._models, self._logw = other._models, other._logw

    def __getnewargs__(self):
        return (self.models, self.log_weights)


# =================================================================================================
# InferredMaskedMix  (src/python_bindings.rs:713-905 ; mix.rs:357-532).  Host numpy ("next").
# =================================================================================================
class InferredMaskedMix:
    def __init__(self, log_posteriors: np.ndarray, inferred: List[InferredMasked], mix: Optional["PPCAMix"] = None):
        self._lp = log_posteriors
        self._inf = inferred
        self._mix = mix

    def __len__(self) -> int:
        return self._lp.shape[0]

    def log_posteriors(self) -> np.ndarray:
        return self._lp.copy() if len(self) else np.zeros((0, 0))

    def posteriors(self) -> np.ndarray:
        return np.exp(self._lp) if len(self) else np.zeros((0, 0))

    def states(self) -> np.ndarray:
        """mix.rs:374-380 weighs the sub-states by the LOG-posterior (reference quirk, reproduced)."""
        if len(self) == 0:
            return np.zeros((0, 0))
        return sum(self._lp[:, j:j + 1] * inf._states for j, inf in enumerate(self._inf))

    def covariances(self) -> List[np.ndarray]:
        mean = self.states()
        post = np.exp(self._lp)
        out = np.zeros_like(self._inf[0]._covs)
        for j, inf in enumerate(self._inf):  # mix.rs:383-394
            dlt = inf._states - mean
            out += post[:, j, None, None] * (inf._covs + np.einsum("na,nb->nab", dlt, dlt))
        return [c for c in out]

    def _parts(self, mix: PPCAMix, dataset: Optional[Dataset]):
        post = np.exp(self._lp)
        x = dataset.numpy() if dataset is not None else None
        parts = []
        for inf, m in zip(self._inf, mix._models):
            sm = inf._states @ m._C.T + m._mu
            parts.append(np.where(np.isfinite(x), x, sm) if x is not None else sm)
        return post, parts

    def smoothed(self, mix: PPCAMix) -> Dataset:
        post, parts = self._parts(mix, None)
        return Dataset(sum(post[:, j:j + 1] * p for j, p in enumerate(parts)))  # mix.rs:397-404

    def extrapolated(self, mix: PPCAMix, dataset: Dataset) -> Dataset:
        post, parts = self._parts(mix, dataset)
        return Dataset(sum(post[:, j:j + 1] * p for j, p in enumerate(parts)))  # mix.rs:407-414

    def _mix_cov_full(self, mix: PPCAMix, dataset: Optional[Dataset]) -> List[np.ndarray]:
        """sum_j p_nj (cov_nj + (m_nj - mean_n)(m_nj - mean_n)^T) (mix.rs:422-437, 466-481): the component matrices come
        from the device (ppca_b200_covariance_full); both the smoothed and the extrapolated form use the SMOOTHED
        component covariance, as the reference does (mix.rs:474)."""
        post, parts = self._parts(mix, dataset)
        mean = sum(post[:, j:j + 1] * p for j, p in enumerate(parts))
        acc = None
        for j, (inf, m) in enumerate(zip(self._inf, mix._models)):
            dlt = parts[j] - mean
            term = post[:, j, None, None] * (inf._cov_full(m, None) + dlt[:, :, None] * dlt[:, None, :])
            acc = term if acc is None else acc + term
        return list(acc)

    def smoothed_covariances(self, mix: PPCAMix) -> List[np.ndarray]:
        return self._mix_cov_full(mix, None)

    def smoothed_covariances_diagonal(self, mix: PPCAMix) -> Dataset:
        post, parts = self._parts(mix, None)
        mean = sum(post[:, j:j + 1] * p for j, p in enumerate(parts))
        acc = 0.0
        for j, (inf, m) in enumerate(zip(self._inf, mix._models)):  # mix.rs:445-458
            acc = acc + post[:, j:j + 1] * (inf._smoothed_cov_diag(m) + (parts[j] - mean) ** 2)
        return Dataset(acc)

    def extrapolated_covariances(self, mix: PPCAMix, dataset: Dataset) -> List[np.ndarray]:
        return self._mix_cov_full(mix, dataset)

    def extrapolated_covariances_diagonal(self, mix: PPCAMix, dataset: Dataset) -> Dataset:
        post, parts = self._parts(mix, dataset)
        mean = sum(post[:, j:j + 1] * p for j, p in enumerate(parts))
        x = dataset.numpy()
        acc = 0.0
        for j, (inf, m) in enumerate(zip(self._inf, mix._models)):  # mix.rs:485-501
            diag = np.where(np.isfinite(x), 0.0, inf._smoothed_cov_diag(m))
            acc = acc + post[:, j:j + 1] * (diag + (parts[j] - mean) ** 2)
        return Dataset(acc)

    def posterior_sampler(self) -> "PosteriorSamplerMix":
        return PosteriorSamplerMix(np.exp(self._lp), [inf.posterior_sampler() for inf in self._inf], self._mix)


class PosteriorSamplerMix:
    """mix.rs:519-532."""

    def __init__(self, posteriors: np.ndarray, samplers: List[PosteriorSampler], mix: Optional[PPCAMix] = None):
        self._post, self._samplers, self._mix = posteriors, samplers, mix

    def sample(self, mix: Optional[PPCAMix] = None, seed: Optional[int] = None) -> Dataset:
        """No arguments, as in the reference (src/python_bindings.rs:895): the samplers carry their models.  Component
        drawn from each row's posterior (WeightedIndex, mix.rs:505-509), then that component's posterior sampler
        (mix.rs:524-531), on the device."""
        mix = mix or self._mix
        if mix is None:
            raise ValueError("a PPCAMix is needed to sample outputs")
        return _posterior_sample(mix._models, self._post, [s_._states for s_ in self._samplers],
                                 [s_._covs for s_ in self._samplers], seed)
