"""bincode 1.3 (little endian, fixed-width ints, u64 lengths) layouts of the reference's serde structs.

Mirrors what `bincode::serialize` produces for (src/python_bindings.rs:66-79,388-401,571-584):
  Dataset     { data: Arc<Vec<MaskedSample>>, weights: Vec<f64> }                  dataset.rs:92-100
  MaskedSample{ data: DVector<f64>, mask: Mask(BitVec<u32>) }                     dataset.rs:10-14, utils.rs:27-28
  PPCAModel   ( Arc<{ output_covariance: { isotropic_noise: f64, transform: DMatrix<f64> }, mean: DVector }> )
                                                                                   ppca_model.rs:18-40
  PPCAMix     ( Arc<{ output_size: usize, models: Vec<PPCAModel>, log_weights: DVector<f64> }> )   mix.rs:27-47
nalgebra's VecStorage serialises as (Vec<T>, nrows, ncols) with column-major data; a Dyn dimension is a
u64 and a Const dimension is zero bytes; bit-vec's BitVec is { storage: Vec<u32>, nbits: usize }.
Byte compatibility with the Rust crate is BELIEVED, not verified: there is no Rust toolchain in this image.
"""
from __future__ import annotations

import struct
from typing import List, Tuple

import numpy as np


class Reader:
    def __init__(self, data: bytes):
        self.b = memoryview(data)
        self.pos = 0

    def u64(self) -> int:
        if self.pos + 8 > len(self.b):
            raise ValueError("io error: unexpected end of file")
        (v,) = struct.unpack_from("<Q", self.b, self.pos)
        self.pos += 8
        return v

    def f64(self) -> float:
        if self.pos + 8 > len(self.b):
            raise ValueError("io error: unexpected end of file")
        (v,) = struct.unpack_from("<d", self.b, self.pos)
        self.pos += 8
        return v

    def array(self, dtype, count: int) -> np.ndarray:
        nbytes = np.dtype(dtype).itemsize * count
        if self.pos + nbytes > len(self.b):
            raise ValueError("io error: unexpected end of file")
        out = np.frombuffer(self.b, dtype=dtype, count=count, offset=self.pos).copy()
        self.pos += nbytes
        return out

    def dvector(self) -> np.ndarray:
        n = self.u64()
        v = self.array("<f8", n)
        nrows = self.u64()
        if nrows != n:
            raise ValueError("invalid DVector: length does not match nrows")
        return v

    def dmatrix(self) -> np.ndarray:
        n = self.u64()
        v = self.array("<f8", n)
        nrows, ncols = self.u64(), self.u64()
        if nrows * ncols != n:
            raise ValueError("invalid DMatrix: length does not match dimensions")
        return np.ascontiguousarray(v.reshape(ncols, nrows).T)  # column-major on the wire


def w_u64(v: int) -> bytes:
    return struct.pack("<Q", int(v))


def w_f64(v: float) -> bytes:
    return struct.pack("<d", float(v))


def w_dvector(v: np.ndarray) -> bytes:
    v = np.ascontiguousarray(v, dtype="<f8").reshape(-1)
    return w_u64(v.size) + v.tobytes() + w_u64(v.size)


def w_dmatrix(m: np.ndarray) -> bytes:
    m = np.asarray(m, dtype="<f8")
    return w_u64(m.size) + np.asfortranarray(m).tobytes(order="F") + w_u64(m.shape[0]) + w_u64(m.shape[1])


# ---- models -------------------------------------------------------------------------------------
def dump_model(sigma: float, transform: np.ndarray, mean: np.ndarray) -> bytes:
    return w_f64(sigma) + w_dmatrix(transform) + w_dvector(mean)


def read_model(r: Reader) -> Tuple[float, np.ndarray, np.ndarray]:
    sigma = r.f64()
    transform = r.dmatrix()
    mean = r.dvector()
    return sigma, transform, mean


def load_model(data: bytes) -> Tuple[float, np.ndarray, np.ndarray]:
    return read_model(Reader(data))


def dump_mix(output_size: int, models: List[Tuple[float, np.ndarray, np.ndarray]], log_weights: np.ndarray) -> bytes:
    out = [w_u64(output_size), w_u64(len(models))]
    for sigma, transform, mean in models:
        out.append(dump_model(sigma, transform, mean))
    out.append(w_dvector(log_weights))
    return b"".join(out)


def load_mix(data: bytes):
    r = Reader(data)
    output_size = r.u64()
    n = r.u64()
    models = [read_model(r) for _ in range(n)]
    log_weights = r.dvector()
    return output_size, models, log_weights


# ---- datasets -----------------------------------------------------------------------------------
def dump_dataset(x_nan: np.ndarray, weights: np.ndarray) -> bytes:
    """x_nan: n x d with NaN at masked slots (Dataset.numpy()).  Masked slots are written as NaN."""
    x = np.ascontiguousarray(x_nan, dtype="<f8")
    n, d = x.shape
    nblocks = (d + 31) // 32
    mask = np.isfinite(x)
    padded = np.zeros((n, nblocks * 32), dtype=bool)
    padded[:, :d] = mask
    # bit-vec: bit i lives in block i / 32 at position i % 32 (LSB first)
    words = np.packbits(padded.reshape(n, nblocks, 32), axis=2, bitorder="little").view("<u4").reshape(n, nblocks)
    parts = [w_u64(n)]
    head = w_u64(d)
    blocks_head = w_u64(nblocks)
    for i in range(n):
        parts.append(head + x[i].tobytes() + head + blocks_head + words[i].tobytes() + head)
    parts.append(w_u64(n) + np.ascontiguousarray(weights, dtype="<f8").tobytes())
    return b"".join(parts)


def load_dataset(data: bytes) -> Tuple[np.ndarray, np.ndarray]:
    r = Reader(data)
    n = r.u64()
    rows = []
    d = None
    for _ in range(n):
        v = r.dvector()
        nblocks = r.u64()
        words = r.array("<u4", nblocks)
        nbits = r.u64()
        if nbits != v.size:
            raise ValueError("invalid MaskedSample: mask length does not match data length")
        bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:nbits].astype(bool)
        row = v.copy()
        row[~bits] = np.nan
        if d is None:
            d = v.size
        rows.append(row)
    nw = r.u64()
    weights = r.array("<f8", nw)
    if nw != n:
        raise ValueError("invalid Dataset: weights length does not match data length")
    x = np.stack(rows) if rows else np.zeros((0, 0))
    return x, weights
