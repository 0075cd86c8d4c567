#!/usr/bin/env python
"""bench.py — EM samples*iterations/s of the B200 PPCA engine (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|...]

A "step" is what the reference trainer does per loop (python/ppca_rs/__init__.py:49-65): the log-likelihood
of the current model plus one EM `iterate` over the whole (per-GPU resident) dataset.  The engine returns the
log-likelihood as a by-product of the E-step, so one step is one pass.

N > 1 is launched by the driver with torch.distributed.run, one rank per GPU.  Samples are sharded across
ranks (weak scaling: every rank holds the same number of rows); each step every rank accumulates its local
statistics buffer, ONE NCCL all-reduce sums it, and every rank finishes the M-step.

`--impl reference` times the CPU restatement of the reference's algorithm (oracle/, kind "port": the Rust crate
cannot be built in this image) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> dict(n per GPU, d, k, p_missing, components, k_true)
WORKLOADS = {
    # configs[0] examples/toy_model.py
    "c1": dict(n=100, d=3, k=2, p=0.2, m=1, k_true=2, desc="toy d=3 k=2 N=100"),
    # configs[1] examples/big_toy_model.py scale, the single-GPU headline
    "c2": dict(n=1_000_000, d=200, k=16, p=0.2, m=1, k_true=16, desc="N=1M d=200 k=16 20% missing"),
    # configs[2] large PPCA: N=100M does not fit (1.64 TB); per-GPU resident shard of 4M rows (65.5 GB)
    "c3": dict(n=4_000_000, d=2048, k=64, p=0.3, m=1, k_true=64, desc="d=2048 k=64 30% missing, 4M-row shard per GPU of the N=100M job"),
    "c3s": dict(n=500_000, d=2048, k=64, p=0.3, m=1, k_true=64, desc="d=2048 k=64 30% missing, 0.5M-row shard"),
    # configs[3] PPCAMix, 32 components
    "c4": dict(n=2_000_000, d=512, k=32, p=0.25, m=32, k_true=32, desc="PPCAMix M=32 d=512 k=32 25% missing, 2M-row shard per GPU"),
    "c4s": dict(n=200_000, d=512, k=32, p=0.25, m=4, k_true=32, desc="PPCAMix M=4 d=512 k=32 25% missing, 0.2M rows"),
    # configs[4] inference only: extrapolate + llks with a trained model, 2M-row shard per GPU of the N=500M job
    "c5": dict(n=2_000_000, d=1024, k=48, p=0.3, m=1, k_true=48, inference=True,
               desc="inference (extrapolate + llks) d=1024 k=48 30% missing, 2M-row shard per GPU of the N=500M job"),
}
SEED = 20240531


def flops_per_sample_iter(d, k):
    """F_iter = 2 d k (k+1) + 6 d k + k^3 (symmetric k x k counted once)."""
    return 2 * d * k * (k + 1) + 6 * d * k + k ** 3


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        # NVML is initialised here, outside the timed region (nvmlInit alone takes tens of ms - a 40 ms region would
        # otherwise be over before the first sample)
        self._nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv = nv
            self._h = nv.nvmlDeviceGetHandleByIndex(index)
            self._max = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    _BITS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def _sample_nvml(self):
        nv = self._nv
        sm = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        self.samples.append([str(sm), str(self._max)] + ["Active" if (r & bit) else "Not Active" for _, bit in self._BITS])

    def _run(self):
        if self._nv is not None:
            try:
                while not self._stop.is_set():
                    self._sample_nvml()
                    self._stop.wait(0.003)
                return
            except Exception:
                pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)
        return False

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for name, col in (("hw_slowdown", 2), ("hw_thermal_slowdown", 3), ("sw_thermal_slowdown", 4), ("sw_power_cap", 5)):
            if any(s[col].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def fp64_peak():
    """FP64 tensor (DMMA) peak for the roofline.  MEASURED_PEAKS.json only has HBM GB/s and bf16 TFLOP/s and
    tcgen05 has no f64 kind, so the denominator is our own register-resident DMMA microbenchmark
    (tools/fp64_peak.cu), run on this box when the binary is present, else the committed round-1 measurement."""
    exe = os.path.join(ROOT, "build", "fp64_peak")
    try:
        if os.path.exists(exe):
            out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
            j = json.loads(out)
            return j["dmma_tflops"], j.get("dmma_tflops_sustained"), "tools/fp64_peak.cu on this box (burst)"
    except Exception:
        pass
    with open(os.path.join(ROOT, "profiles", "r01_fp64_peak.json")) as f:
        j = json.load(f)
    return j["dmma_tflops"], j.get("dmma_tflops_sustained"), "profiles/r01_fp64_peak.json (tools/fp64_peak.cu, round 1)"


def measured_bf16_peak():
    """Dense bf16 TFLOP/s from the driver-written MEASURED_PEAKS.json, else the profiling guide's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops"]), "MEASURED_PEAKS.json"
    except Exception:
        return 1590.0, "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"


def int8_tc_peak():
    """Dense int8 tcgen05 peak for the roofline of the default contraction path.  MEASURED_PEAKS.json has no int8
    figure, so the denominator is our own issue-rate probe (tools/utc_peak.cu: back-to-back tcgen05.mma.kind::i8,
    M=128 N=192 K=32, one CTA per SM), run on this box when the binary is present, else the committed measurement."""
    exe = os.path.join(ROOT, "build", "utc_peak")
    try:
        if os.path.exists(exe):
            out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
            return float(json.loads(out)["utcimma_ts_n192_tops"]), "tools/utc_peak.cu on this box (tcgen05.mma.kind::i8 M128 N192 K32)"
    except Exception:
        pass
    with open(os.path.join(ROOT, "profiles", "r01_utc_i8_peak.json")) as f:
        return float(json.load(f)["utcimma_ts_n192_tops"]), "profiles/r01_utc_i8_peak.json (tools/utc_peak.cu, round 1)"


def measured_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture
    of one chunk of this workload (profiles/r02_traffic.json); None when that workload was not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        wl = {"c3": "c3s"}.get(workload, workload)
        for name, v in t[wl]["kernels"].items():
            if name.startswith(kernel):
                return v["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)"


def ctx_chunk(ctx, n, d, k, slices=None):
    """Samples per chunk the engine picks automatically (mirrors pick_chunk in csrc/api.cu)."""
    slices = slices or int(os.environ.get("PPCA_B200_SLICES", "6"))
    kk = k * (k + 1) // 2
    kkp, kp = (kk + 7) // 8 * 8, (k + 7) // 8 * 8
    per_row = (kkp + 2 * kp + 4) * 8 + kkp * slices
    cap = max(148 * 128, min((8 << 30) // per_row, 2 << 20))
    n_pad = (n + 255) // 256 * 256
    nchunks = -(-n_pad // cap)
    chunk = -(-n_pad // nchunks)
    return min((chunk + 255) // 256 * 256, n_pad)


def init_params(d, k, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((d, k)), np.zeros(d), 1.0


# ------------------------------------------------------------------------------------------------
# reference arm: CPU restatement on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample_rows(wl):
    # about 10 s of CPU work per step on ~16-24 host cores (measured ~1.6e5 samples*iters/s at c2)
    cost = flops_per_sample_iter(wl["d"], wl["k"]) * max(1, wl["m"])
    return int(max(2_000, min(wl["n"], 2.0e11 / cost)))


def host_sample(wl, rows, seed):
    """Same generator semantics as the device one (ppca_model.rs:164-191), numpy on the host."""
    rng = np.random.default_rng(seed)
    d, kt = wl["d"], wl["k_true"]
    Ct = (rng.random((d, kt)) < 0.1).astype(np.float64)
    X = rng.standard_normal((rows, kt)) @ Ct.T + 0.1 * rng.standard_normal((rows, d))
    X[rng.random((rows, d)) < wl["p"]] = np.nan
    return X


def cpu_step_time(wl, X, steps, warmup, keep=None):
    """Times `steps` CPU steps (llk + iterate) of the restated reference algorithm.  `keep` (a dict) receives the
    log-likelihood of the initial model and the model after the FIRST step, for the parity block of the bench line."""
    from oracle import oracle as orc  # bench.py's cpu_baseline / reference leg only
    # the reference fans out on rayon's global pool = every host core; torchrun exports OMP_NUM_THREADS=1, undo that
    orc.set_num_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    d, k, m = wl["d"], wl["k"], wl["m"]
    times = []
    if m == 1:
        C, mu, s = init_params(d, k, SEED + 1000)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            llk = orc.llk(X, None, C, mu, s)            # python/ppca_rs/__init__.py:51
            C, mu, s = orc.iterate(X, None, C, mu, s)  # :61-65
            if it >= warmup:
                times.append(time.perf_counter() - t0)
            if it == 0 and keep is not None:
                keep.update(llk=llk, C=C.copy(), mu=mu.copy(), sigma=s)
    else:
        models = [init_params(d, k, SEED + 1000 + j) for j in range(m)]
        logw = np.full(m, -np.log(m))
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            llk = orc.mix_llk(X, None, models, logw)
            models, logw = orc.mix_iterate(X, None, models, logw)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
            if it == 0 and keep is not None:
                keep.update(llk=llk, models=[(Cj.copy(), muj.copy(), sj) for Cj, muj, sj in models], logw=logw.copy())
    return times, orc.num_threads()


DTYPE_NOTE = {
    "dmma": "f64",
    "tc": "f64; the two masked-Gram contractions sum int8 digit planes exactly (x{T}, tcgen05 kind::i8, every term kept to "
          "8*{T}-2 bits below its column maximum, guarded: repeats at 8 planes / FP64 DMMA when a sample or dimension sits "
          "far below that maximum)",
    "int8": "f64; masked-Gram contractions on int8 digit planes (x{T}, mma.sync IMMA), guarded",
}
PARITY_TOL = 1e-9   # BASELINE.json north_star: per-iteration llk, C, mu, sigma^2 within 1e-9 relative (FP64 path)


def parity_block(pk, ds, wl, rows, keep):
    """The GPU step on the SAME rows and initial model the CPU leg just ran (oracle/ as the checker only)."""
    from oracle import oracle as orc
    d, k, m = wl["d"], wl["k"], wl["m"]
    sub = ds._slice(0, rows)

    def rel(a, b):
        a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))

    X = sub.numpy()
    if m == 1:
        C0, mu0, s0 = init_params(d, k, SEED + 1000)
        new, llk = pk.PPCAModel(s0, C0, mu0)._iterate(sub, None)
        with orc.stable():   # the cancellation-free form of the same algebra (see oracle/ppca_oracle.c header)
            Cs, mus, ss = orc.iterate(X, None, C0, mu0, s0)
        out = {"max_rel_C": rel(new.transform, Cs), "max_rel_mu": rel(new.mean, mus),
               "rel_sigma2": abs(new.isotropic_noise ** 2 - ss ** 2) / ss ** 2,
               "rel_llk": abs(llk - keep["llk"]) / abs(keep["llk"]),
               "vs_reference_order": {"max_rel_C": rel(new.transform, keep["C"]), "max_rel_mu": rel(new.mean, keep["mu"]),
                                      "rel_sigma2": abs(new.isotropic_noise ** 2 - keep["sigma"] ** 2) / keep["sigma"] ** 2}}
    else:
        models = [init_params(d, k, SEED + 1000 + j) for j in range(m)]
        mix = pk.PPCAMix([pk.PPCAModel(sg, Cj, muj) for Cj, muj, sg in models], np.zeros(m))
        new, llk = mix._iterate(sub, None)
        with orc.stable():
            st_models, st_logw = orc.mix_iterate(X, None, models, np.full(m, -np.log(m)))
        out = {"max_rel_C": max(rel(g.transform, w[0]) for g, w in zip(new.models, st_models)),
               "max_rel_mu": max(rel(g.mean, w[1]) for g, w in zip(new.models, st_models)),
               "rel_sigma2": max(abs(g.isotropic_noise ** 2 - w[2] ** 2) / w[2] ** 2 for g, w in zip(new.models, st_models)),
               "max_abs_log_weights": float(np.max(np.abs(new.log_weights - st_logw))),
               "rel_llk": abs(llk - keep["llk"]) / abs(keep["llk"])}
    worst = max(v for kname, v in out.items() if isinstance(v, float))
    out.update(rows=rows, tolerance=PARITY_TOL, ok=bool(worst < PARITY_TOL),
               against="oracle/ppca_oracle.c on the same rows and initial model: C, mu, sigma^2 vs its cancellation-free "
                       "('stable') form, llk vs the reference operation order; vs_reference_order adds the reference's own "
                       "rounding noise (eps (|G|/sigma^2)^2)")
    return out


def run_reference(args, wl, rank):
    if rank != 0:
        return
    rows = cpu_sample_rows(wl)
    X = host_sample(wl, rows, SEED)
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    times, cores = cpu_step_time(wl, X, steps, warmup)
    total = float(sum(times))
    value = rows * len(times) / total
    line = {
        "impl": "reference", "metric": "EM samples*iters/sec", "value": value, "unit": "samples*iters/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # same keys as our arm's config (the workload is the same rows of the same synthetic data set)
        "config": {"workload": f"{args.workload}: {wl['desc']}", "rows_per_gpu": rows, "d": wl["d"], "k": wl["k"],
                   "components": wl["m"],
                   "parallelism": f"OpenMP over samples / dimensions, {cores} host threads (the reference uses rayon the same way)",
                   "l2": "host memory"},
        "steps_note": f"--steps {args.steps} --warmup {args.warmup} requested; the CPU arm runs at most 3 timed "
                      "steps and 1 warm-up (one step takes seconds on the host cores)",
        "cpu_baseline": {"value": value, "unit": "samples*iters/s", "cores": cores, "kind": "port",
                         "sample": f"{rows} rows of the workload, {len(times)} step(s) of llk + iterate "
                                   "(oracle/ppca_oracle.c, OpenMP, restatement of the reference's CPU algorithm; "
                                   "the Rust crate cannot be built in this image)"},
        "e2e": {"value": value, "unit": "samples*iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def timed_steps(torch, ctx, stream, step_fn, steps, warmup, dist, local_rank, sample_clocks=True, profile_inside=True):
    """W untimed + K timed calls of step_fn, CUDA events on the launching stream, barrier + synchronize on both sides,
    max over ranks.  Returns (ms_total, per-family ms, launches, clocks summary).
    profile_inside=False (mixtures): the per-family CUDA events are kept OUT of the timed region — with them the mixture
    chunk loop cannot replay from its CUDA graph — and the family breakdown comes from one extra, untimed step afterwards,
    scaled to `steps`."""
    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step_fn()
    ctx.set_profiling(profile_inside)
    sync_all()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(steps):
            step_fn()
        e1.record(stream)
        sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    if profile_inside:
        fam = ctx.last_profile()
    else:
        ctx.set_profiling(True)
        step_fn()
        sync_all()
        fam = {kname: v * steps for kname, v in ctx.last_profile().items()}
    ctx.set_profiling(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), fam, launches, clocks.summary()


def family_rooflines(fam, ms_total, steps, n, d, k, m, slices, gemm, peak64, peak64_src, tensor_peak, tensor_unit):
    """Every kernel family of the step against the roofline that bounds it (HBM bytes or FP64 / int8 tensor work)."""
    hbm_peak, hbm_src = measured_hbm_peak()
    kk = k * (k + 1) // 2
    kkp, kp = (kk + 7) // 8 * 8, (k + 7) // 8 * 8
    comps = max(1, m)
    fam_bytes = {                                           # algorithmic bytes per sample (per component)
        "proj": 8 * d + d / 8 + 8 * kp,                                   # read x + mask, write y
        "solve": 8 * (2 * kkp + 3 * kp + 4),                              # read G, y; write W, z, w z, scalars
        "slice": (8 + slices) * kkp if gemm != "dmma" else 0,             # read W once, write T digit planes
        "cross_resid": 8 * d + d / 8 + 8 * kp,                            # read x + mask, w z
    }
    # FP64 flops per sample of the families that run on the FP64 pipes (DMMA for proj / cross_resid, DFMA for the solve):
    # with 37 TFLOP/s of FP64 against 6.5 TB/s of HBM the ridge is ~5.7 flop/B, so at k >= 16 the X passes are FP64-bound
    fam_flops = {"proj": 2 * d * k, "cross_resid": 2 * d * k, "solve": 2 * k ** 3}
    families = {}
    for name, ms in fam.items():
        entry = {"ms_per_step": ms / steps, "share_of_step": ms / ms_total if ms_total else None}
        if name in fam_bytes and ms > 0 and fam_bytes[name] > 0:
            gbs = fam_bytes[name] * comps * n * steps / (ms * 1e-3) / 1e9
            tfl = fam_flops.get(name, 0) * comps * n * steps / (ms * 1e-3) / 1e12
            if tfl / peak64 > gbs / hbm_peak:   # the FP64 pipe, not HBM, is the nearer roof
                entry.update(bound="tensor", achieved=tfl, peak=peak64, unit="TFLOP/s", frac=tfl / peak64,
                             peak_source=peak64_src + " (FP64: DMMA rate = DFMA rate)", hbm_gbs=gbs, hbm_frac=gbs / hbm_peak)
            else:
                entry.update(bound="hbm", achieved=gbs, peak=hbm_peak, unit="GB/s", frac=gbs / hbm_peak,
                             peak_source=hbm_src, fp64_tflops=tfl, fp64_frac=tfl / peak64)
        elif name in ("gram", "moment") and ms > 0:
            ops = comps * (2 * d * kk + (2 * d * k if name == "moment" else 0)) * n * steps * (slices if gemm != "dmma" else 1)
            ach = ops / (ms * 1e-3) / 1e12
            entry.update(bound="tensor", achieved=ach, peak=tensor_peak, unit=tensor_unit, frac=ach / tensor_peak)
        families[name] = entry
    return families


def shard_block(torch, pk, pdist, ctx, stream, dist, rank, world, local_rank, args, name, rows, steps, warmup):
    """A second, smaller measurement carried inside the default JSON line: `rows` per GPU of another BASELINE config
    (c3: d=2048 k=64, the north-star shape; c4: PPCAMix M=32), same timing rules, plus the time of its all-reduce."""
    wl = WORKLOADS[name]
    d, k, m = wl["d"], wl["k"], wl["m"]
    ds = pk.Dataset.synthetic(rows, d, wl["k_true"], 0.1, wl["p"], n_components=max(1, m), seed=SEED + 77, ctx=ctx,
                              row_begin=rank * rows)   # every rank holds its own rows of ONE dataset
    if m == 1:
        C, mu, s = init_params(d, k, SEED + 1000)
        state = pdist.ShardedPPCA(ctx, ds, pk.PPCAModel(s, C, mu), group=dist)
    else:
        models = [pk.PPCAModel(sg, Cj, muj) for Cj, muj, sg in (init_params(d, k, SEED + 1000 + j) for j in range(m))]
        state = pdist.ShardedPPCAMix(ctx, ds, pk.PPCAMix(models, np.zeros(m)), group=dist)
    variants0 = ctx.variant_counts()
    ms_total, fam, launches, clocks = timed_steps(torch, ctx, stream, state.step, steps, warmup, dist, local_rank,
                                                  profile_inside=(m == 1))
    variants = {kname: v - variants0.get(kname, 0) for kname, v in ctx.variant_counts().items() if v - variants0.get(kname, 0)}
    block = {"workload": f"{name}: {wl['desc']}", "rows_per_gpu": rows, "d": d, "k": k, "components": m,
             "kernel_variants": variants,
             "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
             "value": rows * world * steps / (ms_total * 1e-3), "unit": "samples*iters/s", "gpu_launches": int(launches),
             "clocks": clocks}
    # the statistics all-reduce of this shape, timed alone on the context's stream (NCCL over NVLink)
    kk = k * (k + 1) // 2
    stats_doubles = max(1, m) * (d * ((kk + 7) // 8 * 8) + d * ((k + 7) // 8 * 8) + 2 * d + 8)
    block["allreduce_bytes"] = stats_doubles * 8
    if dist is not None:
        buf = torch.zeros(stats_doubles, dtype=torch.float64, device="cuda")
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for _ in range(3):
            ctx.comm_allreduce(buf.data_ptr(), stats_doubles)
        torch.cuda.synchronize()
        dist.barrier()
        ev[0].record(stream)
        for _ in range(10):
            ctx.comm_allreduce(buf.data_ptr(), stats_doubles)
        ev[1].record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / 10.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        block["comm_ms_per_step"] = float(t.item())
        block["comm_gb_per_s"] = stats_doubles * 8 / (block["comm_ms_per_step"] * 1e-3) / 1e9
    else:
        block["comm_ms_per_step"] = 0.0
    if rank == 0:
        peak64, _, peak64_src = fp64_peak()
        if args.gemm == "tc":
            tpeak, _ = int8_tc_peak()
            tunit = "TOP/s"
        elif args.gemm == "dmma":
            tpeak, tunit = peak64, "TFLOP/s"
        else:
            with open(os.path.join(ROOT, "profiles", "r01_imma_peak.json")) as f:
                tpeak, tunit = json.load(f)["imma_s8_tops"], "TOP/s"
        fams = family_rooflines(fam, ms_total, steps, rows, d, k, m, args.slices, args.gemm, peak64, peak64_src, tpeak, tunit)
        block["families"] = fams
        bit_ms = fam.get("gram", 0.0) + fam.get("moment", 0.0)
        if bit_ms > 0:
            fl = max(1, m) * (2 * (2 * d * kk) + 2 * d * k) * rows * steps   # Mask Ksym, Mask^T W, Mask^T (w Z)
            ops = fl * (args.slices if args.gemm != "dmma" else 1)
            ach = ops / (bit_ms * 1e-3) / 1e12
            block["contraction"] = {"achieved": ach, "peak": tpeak, "unit": tunit, "frac": ach / tpeak,
                                    "fp64_equivalent_tflops": fl / (bit_ms * 1e-3) / 1e12,
                                    "fp64_dmma_peak_tflops": peak64, "share_of_step": bit_ms / ms_total}
    del state, ds
    torch.cuda.empty_cache()
    return block


def full_job_block(torch, pk, ctx, stream, dist, rank, world, local_rank, args, rows, steps, warmup):
    """BASELINE configs[2] as specified — N = 100 M rows x d = 2048, k = 64, 30 % missing (1.6 TB: fits no set of 8 GPUs) —
    run out of core: every rank owns `rows` consecutive rows of the job (12.5 M = 100 M / 8 by default, so 8 GPUs cover
    the whole job and fewer GPUs cover world/8 of it) and REGENERATES each chunk on the device inside the timed step
    (pk.GeneratedDataset -> ppca_b200_iterate_generated), then one all-reduce of the 35 MB statistics.  The generator's
    time is inside the step (it stands where a storage or network read would)."""
    wl = WORKLOADS["c3"]
    d, k = wl["d"], wl["k"]
    ds = pk.GeneratedDataset(rows, d, wl["k_true"], 0.1, wl["p"], seed=SEED + 501, row_begin=rank * rows, ctx=ctx)
    if dist is not None:
        from ppca_rs_b200 import distributed as pdist
        pdist.native_comm_init(ctx, dist)          # no-op when an earlier block already built the communicator
    C, mu, s = init_params(d, k, SEED + 1000)
    state = {"model": pk.PPCAModel(s, C, mu), "llk": None}

    def step():
        state["model"], state["llk"] = state["model"]._iterate(ds, None, sharded=dist is not None)

    ms_total, fam, launches, clocks = timed_steps(torch, ctx, stream, step, steps, warmup, dist, local_rank)
    return {"workload": "c3 as specified: N=100M d=2048 k=64 30% missing, out of core (chunks regenerated on the device "
                        "inside the step), rows_per_gpu consecutive rows per rank",
            "rows_per_gpu": rows, "rows_total": rows * world, "fraction_of_the_100M_job": rows * world / 100e6,
            "d": d, "k": k, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
            "value": rows * world * steps / (ms_total * 1e-3), "unit": "samples*iters/s", "gpu_launches": int(launches),
            "generator_and_ingest_ms_per_step": (ms_total - sum(fam.values())) / steps,  # the step minus the EM kernel families
            "llk_per_sample_after": state["llk"] / (rows * world) if state["llk"] else None,
            "family_ms_per_step": {kname: v / steps for kname, v in fam.items()}, "clocks": clocks}


def run_ours(args, wl, rank, world, local_rank):
    import torch
    import ppca_rs_b200 as pk
    from ppca_rs_b200 import _native as nat
    from ppca_rs_b200 import distributed as pdist

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream(device=local_rank)  # the default stream's handle is NULL; use a real one
    torch.cuda.set_stream(stream)
    ctx = pk.Context(local_rank, stream.cuda_stream)
    pk.set_context(ctx)
    if args.chunk:
        ctx.set_chunk(args.chunk)
    ctx.set_gemm(args.gemm, args.slices)

    n, d, k, m = wl["n"], wl["d"], wl["k"], wl["m"]
    if args.rows:
        n = args.rows
    ds = pk.Dataset.synthetic(n, d, wl["k_true"], 0.1, wl["p"], n_components=max(1, m), seed=SEED, ctx=ctx, row_begin=rank * n)

    if m == 1:
        C, mu, s = init_params(d, k, SEED + 1000)
        state = pdist.ShardedPPCA(ctx, ds, pk.PPCAModel(s, C, mu), group=dist)
    else:
        models = [pk.PPCAModel(sg, Cj, muj) for Cj, muj, sg in (init_params(d, k, SEED + 1000 + j) for j in range(m))]
        state = pdist.ShardedPPCAMix(ctx, ds, pk.PPCAMix(models, np.zeros(m)), group=dist)

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    variants0 = ctx.variant_counts()
    ms_total, fam, launches, clocks_summary = timed_steps(torch, ctx, stream, state.step, args.steps, args.warmup, dist,
                                                          local_rank, profile_inside=(m == 1))
    variants = {kname: v - variants0.get(kname, 0) for kname, v in ctx.variant_counts().items()
                if v - variants0.get(kname, 0)}
    n_total = n * world
    value = n_total * args.steps / (ms_total * 1e-3)

    # ---- strong scaling: the SAME total work (the N=1 row count) split over the ranks -----------------------------
    strong = None
    if m == 1 and not args.no_blocks:
        n_strong = max(256, n // world)
        ds_s = pk.Dataset.synthetic(n_strong, d, wl["k_true"], 0.1, wl["p"], seed=SEED + 500, ctx=ctx, row_begin=rank * n_strong)
        C, mu, s = init_params(d, k, SEED + 1000)
        st_s = pdist.ShardedPPCA(ctx, ds_s, pk.PPCAModel(s, C, mu), group=dist)
        ms_s, _, _, _ = timed_steps(torch, ctx, stream, st_s.step, args.steps, 3, dist, local_rank, sample_clocks=False)
        strong = {"scaling": "strong", "total_rows": n_strong * world, "rows_per_gpu": n_strong,
                  "ms_per_step": ms_s / args.steps, "value": n_strong * world * args.steps / (ms_s * 1e-3),
                  "unit": "samples*iters/s"}
        del st_s, ds_s

    # ---- the other BASELINE shapes, small shards, inside the same line (same timing rules) -------------------------
    blocks = {}
    if not args.no_blocks and args.workload == "c2":
        blocks["c3_shard"] = shard_block(torch, pk, pdist, ctx, stream, dist, rank, world, local_rank, args, "c3",
                                         args.c3_rows, 5, 3)
        blocks["c4_shard"] = shard_block(torch, pk, pdist, ctx, stream, dist, rank, world, local_rank, args, "c4",
                                         args.c4_rows, 3, 3)
        if args.c3_full_rows > 0:
            blocks["c3_full"] = full_job_block(torch, pk, ctx, stream, dist, rank, world, local_rank, args,
                                               args.c3_full_rows, 1, 3)

    # ---- end to end: HOST buffers in, host model out, every step -----------------------------------------------
    # The samples start in page-locked host memory and cross the bus inside the timed region on every step
    # (single model: ppca_b200_iterate_host / ppca_b200_em_stats_host stream them block by block, H2D overlapped with
    # the kernels; mixtures: Dataset(ndarray) upload + PPCAMix.iterate), the new model is read back to host numpy.
    n_e2e = int(min(n, max(4096, (8 << 30) // (8 * d))))      # bound the pinned host copy to 8 GiB per rank
    Xh = np.empty((n_e2e, d))
    nat.check(nat.lib().ppca_b200_dataset_to_host(ctx.handle, ds._h, 0, n_e2e, nat.dptr(Xh)))
    e2e_steps = max(1, min(args.steps, 10))
    if m == 1:
        C0, mu0, s0 = init_params(d, k, SEED + 1000)
        # compact host format (ppca_b200_pack_host, built once like the reference's own numpy -> Rust copy): only the
        # observed values, the row offsets and the mask words cross the bus each step
        host = pk.HostDataset(Xh, pin=True, ctx=ctx, packed=not args.e2e_plain)
        st8 = pdist.ShardedPPCA(ctx, host, pk.PPCAModel(s0, C0, mu0), group=dist)
        if host._packed is not None:
            vals, rowptr, maskw = host._packed
            h2d = int(rowptr[-1]) * 8 + rowptr.nbytes + maskw.nbytes + (d * k + d) * 8
        else:
            h2d = n_e2e * d * 8 + (d * k + d) * 8
        d2h = (d * k + 2 * d + 8) * 8
        step_fn = st8.step
        fn = "ppca_b200_iterate_host" if args.e2e_plain else "ppca_b200_iterate_packed_host"
        call = (f"PPCAModel.iterate(HostDataset) -> {fn}" if world == 1 else
                f"ShardedPPCA.step over HostDataset shards -> {fn}_sharded (statistics, NCCL all-reduce, finish in one call)")
    else:
        mix0 = pk.PPCAMix([pk.PPCAModel(sg, Cj, muj) for Cj, muj, sg in (init_params(d, k, SEED + 1000 + j) for j in range(m))], np.zeros(m))
        host = pk.HostDataset(Xh, pin=True, ctx=ctx)            # page-locks Xh; Dataset(Xh) below DMAs from it directly
        box = {"mix": mix0}
        h2d = n_e2e * d * 8 + m * (d * k + d) * 8 * 2
        d2h = m * (d * k + 2 * d + 8) * 8

        def step_fn():
            dsh = pk.Dataset(Xh, _ctx=ctx)                      # numpy -> device every step
            st = pdist.ShardedPPCAMix(ctx, dsh, box["mix"], group=dist)
            st.step()
            box["mix"] = st.mix
        call = "Dataset(ndarray) + ShardedPPCAMix.step (ppca_b200_dataset_from_host, ppca_b200_mix_posteriors, ppca_b200_mix_em_stats, ppca_b200_em_finish)"
    for _ in range(2):
        step_fn()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_fn()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    # the same call with the Dataset left behind its device handle (how the reference's own API holds it)
    resident = None
    if world == 1:
        if m == 1:
            C0, mu0, s0 = init_params(d, k, SEED + 1000)
            model = pk.PPCAModel(s0, C0, mu0)
        else:
            model = mix0
        for _ in range(2):
            model, _ = model._iterate(ds, None)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            model, _ = model._iterate(ds, None)
        torch.cuda.synchronize()
        resident = n * e2e_steps / (time.perf_counter() - t0)
    e2e = {"value": n_e2e * world * e2e_steps / e2e_s, "unit": "samples*iters/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "steps": e2e_steps, "rows_per_gpu": n_e2e, "call": call,
           "h2d_gb_per_s": h2d * e2e_steps / e2e_s / 1e9,
           "resident_handle_value": resident,
           "host_bytes_per_sample": h2d / n_e2e, "dense_bytes_per_sample": d * 8,
           "note": "samples in page-locked HOST memory cross the bus every step inside the timed region (wall clock, "
                   "barrier + synchronize on both sides, max over ranks); resident_handle_value = same public call with "
                   "the Dataset kept behind its device handle, as the reference API holds it (src/python_bindings.rs:28-30)"}
    del host, Xh

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel family: the masked-Gram contraction (E-step + M-step launches) ------
    peak64, peak64_sustained, peak64_src = fp64_peak()
    kk = k * (k + 1) // 2
    bit_ms = fam.get("gram", 0.0) + fam.get("moment", 0.0)
    # one E-step and one M-step contraction per model (the mixture pass shares the E-step between the posteriors and the
    # weighted statistics)
    contractions = 2 * max(1, m)
    # algorithmic FP64 flops: Mask * Ksym (E-step), Mask^T * W and Mask^T * (w Z) (M-step), per model
    flops_bit = max(1, m) * (2 * (2 * d * kk) + 2 * d * k) * n * args.steps
    fp64_equiv = flops_bit / (bit_ms * 1e-3) / 1e12 if bit_ms > 0 else None
    launches_bit = contractions * args.steps * max(1, -(-n // max(1, ctx_chunk(ctx, n, d, k))))
    n_chunk = min(ctx_chunk(ctx, n, d, k), n)
    # one launch = one chunk: the packed mask in, G (E-step, 8 B per entry) or the W digit planes (M-step, `slices` B)
    alg_bytes_launch = n_chunk * ((d + 31) // 32) * 4 + n_chunk * ((kk + 7) // 8 * 8) * (8 + args.slices) / 2.0
    common = {
        "bound": "tensor", "traffic": measured_traffic(args.workload, "tbitgemm_atm_kernel") if args.gemm == "tc" else None,
        "traffic_note": "dram__bytes_read + dram__bytes_write per launch (E- and M-step launches averaged) from the committed "
                        "ncu --set full capture (profiles/r02_traffic.json)",
        "algorithmic_bytes_per_launch": alg_bytes_launch,
        "share_of_step": bit_ms / ms_total if ms_total else None,
        "avg_launch_ms": bit_ms / launches_bit if launches_bit else None,
        "family_ms_per_step": {kname: v / args.steps for kname, v in fam.items()},
        "fp64_equivalent_tflops": fp64_equiv, "fp64_dmma_peak_tflops": peak64, "fp64_peak_source": peak64_src,
        "fp64_equivalent_frac": (fp64_equiv / peak64) if fp64_equiv else None,
        "whole_step_fp64_equivalent_frac":
            flops_per_sample_iter(d, k) * max(1, m) * n * args.steps / (ms_total * 1e-3) / 1e12 / peak64,
    }
    if args.gemm == "dmma":
        roofline = dict(common, kernel="bitgemm_kernel (mma.sync DMMA; E-step + M-step launches)", achieved=fp64_equiv,
                        peak=peak64, unit="TFLOP/s", frac=common["fp64_equivalent_frac"], peak_source=peak64_src,
                        peak_sustained=peak64_sustained)
    else:
        tops = flops_bit * args.slices / (bit_ms * 1e-3) / 1e12 if bit_ms > 0 else None   # int8 ops = T digit planes
        if args.gemm == "tc":
            peak, psrc = int8_tc_peak()
            bf16, src = measured_bf16_peak()
            psrc += f"; for scale: 2 x bf16_tflops of {src} = {2.0 * bf16:.0f}"
            kern = f"tbitgemm_kernel<{args.slices}> (tcgen05.mma.kind::i8, TMEM accumulators; E-step + M-step launches)"
        else:
            with open(os.path.join(ROOT, "profiles", "r01_imma_peak.json")) as f:
                peak = json.load(f)["imma_s8_tops"]
            psrc = "profiles/r01_imma_peak.json (tools/imma_peak.cu, mma.sync.m16n8k32.s8, round 1)"
            kern = f"ibitgemm_kernel<{args.slices}> (mma.sync IMMA; E-step + M-step launches)"
        roofline = dict(common, kernel=kern, achieved=tops, peak=peak, unit="TOP/s", frac=(tops / peak) if tops else None,
                        peak_source=psrc,
                        note=f"exact int8-sliced evaluation: {args.slices} balanced base-256 digit planes (int8) per FP64 operand, "
                             "int32 accumulation, FP64 recombination; achieved counts 2*rows*d*kk*slices int8 ops per launch")

    # ---- every kernel family of the step against the roofline that bounds it -----------------------------------
    families = family_rooflines(fam, ms_total, args.steps, n, d, k, m, args.slices, args.gemm, peak64, peak64_src,
                                roofline["peak"], roofline["unit"])
    roofline["families"] = families
    # the family that takes the largest share of the step, with the roof that bounds it (the top-level entry above is the
    # masked-Gram contraction the north star names; at small k the per-sample solve is the longer kernel)
    timed = {kname: e for kname, e in families.items() if e.get("frac") is not None}
    if timed:
        top = max(timed, key=lambda kname: timed[kname]["ms_per_step"])
        roofline["dominant_by_time"] = dict(timed[top], family=top)

    # ---- CPU baseline on a bounded sample of the same workload (rank 0, N=1 only) ----------------
    cpu, parity = None, None
    if world == 1 and not args.no_cpu:
        rows = min(cpu_sample_rows(wl), n)
        X = np.empty((rows, d))
        nat.check(nat.lib().ppca_b200_dataset_to_host(ctx.handle, ds._h, 0, rows, nat.dptr(X)))
        keep = {}
        times, cores = cpu_step_time(wl, X, 1, 0, keep)
        parity = parity_block(pk, ds, wl, rows, keep)
        cpu = {"value": rows / times[0], "unit": "samples*iters/s", "cores": cores, "kind": "port",
               "sample": f"first {rows} rows of the GPU workload, 1 step of llk + iterate with the CPU restatement "
                         "of the reference's algorithm (oracle/ppca_oracle.c, OpenMP); Rust crate not buildable here"}

    # ---- one-off ingest cost of the public Dataset(ndarray) constructor (host -> device, mask build) ----------
    ingest = None
    if world == 1:
        rows_i = min(n, max(1, int(4e8 // (8 * d))))
        Xh = np.empty((rows_i, d))
        nat.check(nat.lib().ppca_b200_dataset_to_host(ctx.handle, ds._h, 0, rows_i, nat.dptr(Xh)))
        _warm = pk.Dataset(np.ascontiguousarray(Xh[:4096]))  # first use allocates the pinned staging blocks
        del _warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ds_h = pk.Dataset(Xh)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ingest = {"rows": rows_i, "seconds": dt, "gb_per_s": rows_i * d * 8 / dt / 1e9,
                  "note": "pk.Dataset(ndarray) from pageable host memory: H2D copy + mask/transpose kernels; paid once "
                          "per dataset, like the reference's own numpy -> Rust copy (src/python_bindings.rs:41-54)"}
        del ds_h

    line = {
        "metric": "EM samples*iters/sec", "value": value, "unit": "samples*iters/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": DTYPE_NOTE.get(args.gemm, "f64").format(T=args.slices),
        "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "rows_per_gpu": n, "d": d, "k": k,
                   "components": m, "parallelism": f"sample-sharded x{world}, one NCCL all-reduce of the statistics per step "
                                                   "(ppca_b200_iterate_sharded: collective inside the C ABI)",
                   "l2": "inputs larger than L2 (resident X per GPU = %.2f GB)" % (n * d * 8 / 1e9)},
        "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "ingest": ingest,
        "gpu_launches": int(launches), "kernel_variants": variants, "clocks": clocks_summary,
        "strong_scaling": strong, "c3_shard": blocks.get("c3_shard"), "c4_shard": blocks.get("c4_shard"),
        "c3_full": blocks.get("c3_full"),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        print("bench.py: GPU step differs from the oracle by more than %.0e: %s" % (PARITY_TOL, parity), file=sys.stderr)
        sys.exit(3)


# ------------------------------------------------------------------------------------------------
# inference workload (BASELINE configs[4]): extrapolate + llks with a trained model
# ------------------------------------------------------------------------------------------------
def run_inference(args, wl, rank, world, local_rank):
    import ctypes as C
    import torch
    import ppca_rs_b200 as pk
    from ppca_rs_b200 import _native as nat

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream(device=local_rank)
    torch.cuda.set_stream(stream)
    ctx = pk.Context(local_rank, stream.cuda_stream)
    pk.set_context(ctx)
    ctx.set_gemm(args.gemm, args.slices)
    n, d, k = (args.rows or wl["n"]), wl["d"], wl["k"]
    ds = pk.Dataset.synthetic(n, d, wl["k_true"], 0.1, wl["p"], seed=SEED, ctx=ctx, row_begin=rank * n)
    # "a trained model": 10 EM iterations on a subsample, same on every rank
    sub = pk.Dataset.synthetic(min(n, 262_144), d, wl["k_true"], 0.1, wl["p"], seed=SEED, ctx=ctx)
    C0, mu0, s0 = init_params(d, k, SEED + 1000)
    model = pk.PPCAModel(s0, C0, mu0)
    for _ in range(10):
        model = model.iterate(sub)
    del sub

    def sync_all():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    box = {"out": None}

    def step():
        # ONE E-step yields both outputs (ppca_b200_reconstruct); the fully observed output Dataset of the previous
        # step is overwritten in place, the n log-likelihoods go to the host
        ex, ll = model.reconstruct(ds, True, out=box["out"], with_llks=True)
        box["out"] = ex
        return ex, ll

    for _ in range(max(3, args.warmup)):
        step()
    ctx.set_profiling(True)
    sync_all()
    launches0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        sync_all()
    fam = ctx.last_profile()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    ctx.set_profiling(False)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n * world * args.steps / (ms_total * 1e-3)

    # end to end: host samples in, host reconstruction + llks out, through the C ABI (three-stream pipeline)
    n_e2e = int(min(n, max(4096, (4 << 30) // (8 * d))))
    Xh = np.empty((n_e2e, d))
    nat.check(nat.lib().ppca_b200_dataset_to_host(ctx.handle, ds._h, 0, n_e2e, nat.dptr(Xh)))
    out, ll = np.empty((n_e2e, d)), np.empty(n_e2e)
    for a in (Xh, out, ll):
        nat.check(nat.lib().ppca_b200_host_register(C.c_void_p(a.ctypes.data), a.nbytes))

    def e2e_step():
        nat.check(nat.lib().ppca_b200_reconstruct_host(ctx.handle, nat.dptr(Xh), n_e2e, d, k, nat.dptr(model._C),
                                                       nat.dptr(model._mu), model._sigma, 1, nat.dptr(out), nat.dptr(ll)))
    e2e_steps = max(1, min(args.steps, 5))
    e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    fin = np.isfinite(Xh)
    assert np.array_equal(out[fin], Xh[fin]) and np.isfinite(out).all()      # observed slots are bit copies
    for a in (Xh, out, ll):
        nat.lib().ppca_b200_host_unregister(C.c_void_p(a.ctypes.data))
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    hbm_peak, hbm_src = measured_hbm_peak()
    kk = k * (k + 1) // 2
    alg_bytes = 8 * d + d / 8 + 8 * d + 8                    # x + mask in, reconstruction + llk out
    gbs = alg_bytes * n * args.steps / (ms_total * 1e-3) / 1e9
    peak_i8, psrc = int8_tc_peak()
    gram_ms = fam.get("gram", 0.0)
    tops = (2 * d * kk) * args.slices * n * args.steps / (gram_ms * 1e-3) / 1e12 if gram_ms > 0 else None
    line = {
        "metric": "inference samples/sec (extrapolate + llks)", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "rows_per_gpu": n, "d": d, "k": k,
                   "parallelism": f"sample-sharded x{world}, no collective",
                   "l2": "inputs larger than L2 (resident X per GPU = %.2f GB)" % (n * d * 8 / 1e9)},
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                     "traffic": None, "peak_source": hbm_src,
                     "note": "whole step against its streaming bound (read x + mask, write reconstruction + llk = %.0f B "
                             "per sample): the step is NOT HBM-bound in FP64 - the per-sample k x k solve and the masked-Gram "
                             "contraction dominate (one E-step pass serves extrapolate and llks: PPCAModel.reconstruct -> "
                             "ppca_b200_reconstruct), see family_ms_per_step" % alg_bytes,
                     "family_ms_per_step": {kn: v / args.steps for kn, v in fam.items()},
                     "gram_int8_tops": tops, "gram_int8_frac": (tops / peak_i8) if tops else None, "int8_peak_source": psrc},
        "cpu_baseline": None,
        "e2e": {"value": n_e2e * world * e2e_steps / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": n_e2e * d * 8,
                "d2h_bytes_per_step": n_e2e * (d + 1) * 8, "steps": e2e_steps, "rows_per_gpu": n_e2e,
                "call": "ppca_b200_reconstruct_host (HOST x in, HOST extrapolation + llks out; H2D, kernels and D2H on three streams)",
                "h2d_gb_per_s": n_e2e * d * 8 * e2e_steps / e2e_s / 1e9},
        "gpu_launches": int(launches), "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU")
    ap.add_argument("--chunk", type=int, default=0, help="samples per chunk (0 = automatic)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-plain", action="store_true",
                    help="end-to-end leg streams the full n x d matrix instead of the compact host format")
    ap.add_argument("--no-blocks", action="store_true", help="skip the strong-scaling run and the c3/c4 shard blocks")
    ap.add_argument("--c3-rows", type=int, default=524_288, help="rows per GPU of the c3_shard block")
    ap.add_argument("--c4-rows", type=int, default=131_072, help="rows per GPU of the c4_shard block")
    ap.add_argument("--c3-full-rows", type=int, default=12_500_000,
                    help="rows per GPU of the c3_full block (N=100M / 8; chunks regenerated on the device); 0 skips it")
    ap.add_argument("--gemm", default=os.environ.get("PPCA_B200_GEMM", "tc"), choices=["dmma", "int8", "tc"],
                    help="arithmetic path of the masked-Gram contractions (see include/ppca_b200.h)")
    ap.add_argument("--slices", type=int, default=int(os.environ.get("PPCA_B200_SLICES", "6")))
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    if wl.get("inference"):
        run_inference(args, wl, rank, world, local_rank)
        return
    run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
