#!/usr/bin/env python
"""Phase timing of solve_tile_kernel (development build: make -C ppca_rs_b200/csrc NVCC_EXTRA=-DPPCA_SOLVE_TIMING
OUT=../libppca_b200_timing.so): cycles per warp and sample spent loading / in the publishers' panel section / at the panel
barrier / in the rank-1 updates / after the elimination.  Usage: python tools/solve_phase_timing.py [k] [d] [rows]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ppca_rs_b200 import _native as nat  # noqa: E402

nat.LIB_PATH = os.path.join(ROOT, "ppca_rs_b200", "libppca_b200_timing.so")
import numpy as np  # noqa: E402
import ppca_rs_b200 as pk  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
n = int(sys.argv[3]) if len(sys.argv) > 3 else 131072
ds = pk.Dataset.synthetic(n, d, k, 0.1, 0.3, seed=5)
rng = np.random.default_rng(0)
model = pk.PPCAModel(1.0, rng.standard_normal((d, k)), np.zeros(d))
lib = nat.lib()
lib.ppca_b200_debug_solve_timing.argtypes = [C.POINTER(C.c_uint64), C.c_int]
out = (C.c_uint64 * 8)()
model._iterate(ds, None)
lib.ppca_b200_debug_solve_timing(out, 1)
model._iterate(ds, None)
lib.ppca_b200_debug_solve_timing(out, 1)
v = list(out)
names = ["load+gather", "panel section (publishers)", "panel barrier", "rank-1 updates", "after the elimination"]
tot = sum(v[:5])
print(f"k={k} d={d} rows={n}: warp-samples {v[5]}, cycles per warp and sample {tot / max(v[5], 1):.0f}")
for nm, c in zip(names, v[:5]):
    print(f"  {nm:30s} {c / max(v[5], 1):9.0f} cycles  {100 * c / tot:5.1f} %")
