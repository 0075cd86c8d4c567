#!/usr/bin/env python
"""Regenerates profiles/r02_summary.md from the committed round-2 bench JSON lines (profiles/r02_bench_*.json)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
FAMS = ("gram", "proj", "solve", "slice", "moment", "cross_resid", "finish")


def L(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def famrow(fam):
    return " | ".join(f"{fam[k]['ms_per_step']:.2f}" if k in fam else "—" for k in FAMS)


def roofrow(fam):
    out = []
    for k in ("gram", "moment", "proj", "solve", "slice", "cross_resid"):
        e = fam.get(k, {})
        out.append(f"{100 * e['frac']:.0f} % {e['bound']}" if e.get("frac") is not None else "—")
    return " | ".join(out)


def main():
    c2, c2x2, c3s, c4, c5, ref = (L("r02_bench_c2.json"), L("r02_bench_c2_2gpu.json"), L("r02_bench_c3s.json"),
                                  L("r02_bench_c4.json"), L("r02_bench_c5.json"), L("r02_bench_c2_reference_arm.json"))
    r1c2, r1c3s, r1c5 = L("r01_bench_c2.json"), L("r01_bench_c3s.json"), L("r01_bench_c5.json")
    b3, b4 = c2["c3_shard"], c2["c4_shard"]
    par = c2["parity"]
    lines = [
        "# Round 2 — measured results (one B200 unless noted, sm_100a, SM clock 1965 MHz, no throttle reasons)",
        "",
        "All numbers come from `bench.py` on fresh `gpurun` boxes; the JSON lines are committed next to this file",
        "(`r02_bench_*.json`).  ncu evidence: `r02_c2_step_ncu_full_summary.txt`, `r02_c3s_step_ncu_full_summary.txt`,",
        "`r02_*_step_ncu_details.txt` (`--set full`, one chunk each), `r02_launches_c2.csv` / `r02_launches_c3s.csv` (launch lists",
        "of the bench command), `r02_traffic.json` (DRAM bytes per launch), `r02_tbitgemm_sass_mnemonics.txt` (tcgen05 / TMA",
        "opcodes per contraction kernel), `r02_solve_sass_breakdown.md`, `r02_solve_accuracy.md`, `r02_sanitizer_*.log`",
        "(compute-sanitizer memcheck / racecheck, 0 errors / 0 hazards).  Peaks: `r01_fp64_peak.json` (DMMA 37.1 TFLOP/s = DFMA",
        "rate), `r01_utc_i8_peak.json` (tcgen05 int8 4.3 POP/s), `MEASURED_PEAKS.json` (HBM 6551 GB/s).",
        "",
        "## Headline (BASELINE metric: EM samples·iterations/s; one iteration = llk of the current model + `iterate`)",
        "",
        "| workload | rows/GPU | ms/step (round 1) | ms/step | value (dataset resident) | e2e (host samples cross the bus every step) |",
        "|---|---|---|---|---|---|",
        f"| c2: d=200 k=16 20 % missing | 1 000 000 | {r1c2['ms_per_step']:.2f} | {c2['ms_per_step']:.2f} | {c2['value']/1e6:.1f} M/s | "
        f"{c2['e2e']['value']/1e6:.1f} M/s ({c2['e2e']['host_bytes_per_sample']:.0f} B/sample over PCIe at {c2['e2e']['h2d_gb_per_s']:.1f} GB/s; "
        f"round 1: {r1c2['e2e']['value']/1e6:.1f} M/s with the dense 1600 B/sample) |",
    ]
    if c3s:
        lines.append(f"| c3 shard: d=2048 k=64 30 % missing | 500 000 | {r1c3s['ms_per_step']:.1f} | {c3s['ms_per_step']:.1f} | "
                     f"{c3s['value']/1e6:.2f} M/s | {c3s['e2e']['value']/1e6:.2f} M/s (round 1: {r1c3s['e2e']['value']/1e6:.2f}) |")
    lines.append(f"| c3 shard inside the default line (`c3_shard`) | {b3['rows_per_gpu']:,} | — | {b3['ms_per_step']:.1f} | {b3['value']/1e6:.2f} M/s | — |")
    lines.append(f"| c4 shard inside the default line (`c4_shard`): PPCAMix M=32 d=512 k=32 | {b4['rows_per_gpu']:,} | — | {b4['ms_per_step']:.1f} | "
                 f"{b4['value']/1e6:.2f} M samples·iters/s = {32 * b4['value']/1e6:.1f} M component-samples/s "
                 f"(round 1, M=4: 35.6 M component-samples/s with two E-steps per component) | — |")
    bf = c2.get("c3_full")
    if bf:
        lines.append(f"| c3 AS SPECIFIED (`c3_full`): N = 100 M x 2048, k = 64, out of core, chunks regenerated on the device inside the step | "
                     f"{bf['rows_per_gpu']:,} (= N/8) | — | {bf['ms_per_step']:.0f} | {bf['value']/1e6:.2f} M/s per GPU "
                     f"(generator + ingest {bf['generator_and_ingest_ms_per_step']:.0f} ms of the step) | — |")
    if c4:
        n4 = c4["config"]["rows_per_gpu"]
        lines.append(f"| c4: PPCAMix M=32 d=512 k=32 25 % missing | {n4:,} | — | {c4['ms_per_step']:.0f} | {c4['value']/1e6:.2f} M samples·iters/s = "
                     f"{32 * c4['value']/1e6:.1f} M component-samples/s | {c4['e2e']['value']/1e6:.2f} M/s |")
    if c5:
        lines.append(f"| c5 inference (extrapolate + llks, one E-step) d=1024 k=48 | 2 000 000 | {r1c5['ms_per_step']:.0f} | {c5['ms_per_step']:.0f} | "
                     f"{c5['value']/1e6:.2f} M samples/s | {c5['e2e']['value']/1e6:.2f} M samples/s (host in, host out) |")
    lines += ["",
              f"CPU arm (`--impl reference`, oracle port, {ref['cpu_baseline']['cores']} cores): {ref['value']/1e3:.0f} K samples·iters/s at c2 — resident "
              f"{c2['value']/ref['value']:.0f}×, end to end {c2['e2e']['value']/ref['value']:.0f}×." if ref else "",
              "",
              f"Parity at the bench config (GPU step vs the oracle on the same {par['rows']:,} rows, emitted by every run): "
              f"C {par['max_rel_C']:.1e}, μ {par['max_rel_mu']:.1e}, σ² {par['rel_sigma2']:.1e}, llk {par['rel_llk']:.1e} (bar 1e-9).",
              ""]
    if c2x2:
        s2 = c2x2["strong_scaling"]
        lines += [f"2 GPUs (`gpurun --gpus 2`, torchrun, NCCL inside the library): c2 weak {c2x2['value']/1e6:.0f} M/s "
                  f"({c2x2['value']/c2['value']:.2f}× of one GPU), strong (1 M rows total) {s2['value']/1e6:.0f} M/s, e2e "
                  f"{c2x2['e2e']['value']/1e6:.1f} M/s; c3 shard {c2x2['c3_shard']['value']/1e6:.2f} M/s with its 35 MB all-reduce at "
                  f"{c2x2['c3_shard']['comm_ms_per_step']:.3f} ms ({c2x2['c3_shard'].get('comm_gb_per_s', 0):.0f} GB/s); c4 shard {c2x2['c4_shard']['value']/1e6:.2f} M/s "
                  f"({c2x2['c4_shard']['ms_per_step']:.0f} ms; all-reduce of {c2x2['c4_shard']['allreduce_bytes']/1e6:.0f} MB, all 32 components at once, "
                  f"{c2x2['c4_shard']['comm_ms_per_step']:.3f} ms); c3 as specified {c2x2['c3_full']['value']/1e6:.1f} M/s.", ""]
    c8 = L("r02_bench_c2_8gpu.json")
    if c8:
        lines += [f"8 GPUs (`gpurun --gpus 8`, the driver's command `bench.py --gpus 8 --steps 20 --warmup 5`): c2 weak {c8['value']/1e6:.0f} M/s "
                  f"({c8['value']/c2['value']:.2f}x of one GPU, {c8['ms_per_step']:.2f} ms per step), strong (1 M rows total) {c8['strong_scaling']['value']/1e6:.0f} M/s, "
                  f"e2e {c8['e2e']['value']/1e6:.0f} M/s ({c8['e2e']['h2d_gb_per_s']:.1f} GB/s of H2D per GPU: the host, not the engine, is the limit); "
                  f"c3 shard {c8['c3_shard']['value']/1e6:.1f} M/s (35 MB all-reduce {c8['c3_shard']['comm_ms_per_step']:.3f} ms); "
                  f"c4 shard {c8['c4_shard']['value']/1e6:.2f} M/s; **c3 as specified, all N = 100 M rows: {c8['c3_full']['value']/1e6:.1f} M samples·iters/s, "
                  f"{c8['c3_full']['ms_per_step']/1e3:.2f} s per EM iteration**.", ""]
    c4g = L("r02_bench_c2_4gpu.json")
    if c8 and c4g and c2x2:
        rows_ = [(1, c2), (2, c2x2), (4, c4g), (8, c8)]
        lines += ["### Scaling on one box (every line: `bench.py --gpus N` under torchrun, NCCL inside the library)", "",
                  "| GPUs | c2 weak (M/s) | eff. | c2 strong, 1 M rows total (M/s) | c2 e2e (M/s) | H2D per GPU (GB/s) | c3 shard (M/s) | c4 shard (M/s) | c3 as specified (M/s) |",
                  "|---|---|---|---|---|---|---|---|---|"]
        for n_, j_ in rows_:
            ss = j_.get("strong_scaling") or {}
            lines.append(f"| {n_} | {j_['value']/1e6:.0f} | {j_['value']/c2['value']/n_:.3f} | {ss.get('value', 0)/1e6:.0f} | {j_['e2e']['value']/1e6:.1f} | "
                         f"{j_['e2e']['h2d_gb_per_s']:.1f} | {j_['c3_shard']['value']/1e6:.1f} | {j_['c4_shard']['value']/1e6:.2f} | {j_['c3_full']['value']/1e6:.1f} |")
        lines.append("")
    lines += ["## Second half of round 2", "",
              "* Per-sample solve, 32 < k <= 64: register-tiled, panel-blocked sweep (`solve_tile_kernel`): c5 solve 60.2 -> 46.9 ms, c3s 29.7 -> 26.4 ms;",
              "  seven measured variants and the ncu analysis (latency-bound serial chain per pivot, four samples resident per SM) in",
              "  `r02_solve_tile.md`; `tools/rank1_probe.cu`: the rank-1 update pattern alone runs at 87 % of the FP64 peak.",
              "* Mixture EM: the chunk loop (4 500 launches per step at M = 32) replays from a captured CUDA graph: c4 shard 223 -> 99 ms on a",
              "  host whose launch rate bounded the step (`r02_launches_c4.csv`: 49 ms of kernels per 65 536 rows); on fast hosts 106 -> 99 ms.",
              "  The ranks of the bench hold row ranges of ONE synthetic dataset (`Dataset.synthetic(..., row_begin=rank * rows)`); when every",
              "  rank drew its own truth (earlier runs), a rank whose random start left components with next to no responsibility mass",
              "  repeated those components up the precision ladder every step (2 GPUs: 120 repeats, 201-239 ms instead of 103).",
              "* Out-of-core EM over regenerated chunks (`ppca_b200_iterate_generated`): BASELINE configs[2] as specified; the generator",
              "  (xi -> DMMA row GEMM -> noise + mask) went from 5.9 s to 0.31 s per 12.5 M rows.",
              "* Widened into SURVEY §8(f): device ingestion (`__cuda_array_interface__` / DLPack), every sampler and the full covariances on the device.",
              "* compute-sanitizer over all of it: memcheck 0 errors, racecheck 0 hazards (`r02_sanitizer_*.log`); 170 GPU tests (`r02_pytest_gpu.log`).", "",
              "## Kernel families, ms per step (CUDA events on the launching stream, inside the timed region)", "",
              "| workload | gram (E) | proj | solve | slice | moment (M) | cross | finish |", "|---|---|---|---|---|---|---|---|",
              f"| c2 round 1 | {famrow(r1c2['roofline']['families'])} |",
              f"| c2 | {famrow(c2['roofline']['families'])} |",
              f"| c3s round 1 | {famrow(r1c3s['roofline']['families'])} |"]
    if c3s:
        lines.append(f"| c3s | {famrow(c3s['roofline']['families'])} |")
    lines.append(f"| c3_shard block | {famrow(b3['families'])} |")
    lines.append(f"| c4_shard block (32 components) | {famrow(b4['families'])} |")
    lines += ["", "Fraction of the nearer roof per family (tensor = tcgen05 int8 peak for gram/moment, FP64 37.1 TFLOP/s for the others;",
              "hbm = 6551 GB/s):", "", "| workload | gram | moment | proj | solve | slice | cross |", "|---|---|---|---|---|---|---|",
              f"| c2 | {roofrow(c2['roofline']['families'])} |"]
    if c3s:
        lines.append(f"| c3s | {roofrow(c3s['roofline']['families'])} |")
    lines.append(f"| c4_shard | {roofrow(b4['families'])} |")
    cz = b3.get("contraction", {})
    lines += ["", "## Masked-Gram contraction at the north-star shape (c3: d=2048, k=64; `tbitgemm_atm2_kernel<6>`)", "",
              f"{cz.get('achieved', 0):.0f} TOP/s = {100 * cz.get('frac', 0):.1f} % of the measured tcgen05 int8 peak = "
              f"{cz.get('fp64_equivalent_tflops', 0):.0f} FP64-equivalent TFLOP/s = "
              f"{cz.get('fp64_equivalent_tflops', 0) / cz.get('fp64_dmma_peak_tflops', 37.1):.1f}× the FP64 DMMA roofline; ncu: tensor pipe active "
              "64-69 % of cycles on these launches (`r02_c3s_step_ncu_full_summary.txt`).  c2 (k=16, d=200) is not tensor-bound: a tile is two K",
              "steps and 136 output columns, so the contraction is bounded by the FP64 recombination epilogue and the G write "
              f"({100 * c2['roofline']['frac']:.1f} % of the int8 peak).", ""]
    open(os.path.join(P, "r02_summary.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
