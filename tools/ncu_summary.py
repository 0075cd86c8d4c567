#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers we track."""
import csv, subprocess, sys, re

KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_B", "launch__shared_mem_per_block_dynamic"),
    ("occ_warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("dmma_pipe_pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("inst", "smsp__inst_executed.sum"),
]
STALLS = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio$|smsp__average_warp_latency_issue_stalled_(\w+).ratio$")


def to_base(val, unit):
    v = float(val.replace(",", "")) if val not in ("", "n/a") else float("nan")
    u = unit.lower()
    scale = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "byte": 1e-6, "kbyte": 1e-3, "mbyte": 1, "gbyte": 1e3}
    return v * scale.get(u, 1)


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]][:90])
        for name, key in KEYS:
            if key in idx:
                print(f"   {name:18s} {to_base(r[idx[key]], units[idx[key]]):14.3f}")
        stalls = []
        for h, i in idx.items():
            m = STALLS.match(h)
            if m and r[i] not in ("", "n/a"):
                stalls.append((float(r[i].replace(",", "")), m.group(1) or m.group(2)))
        stalls.sort(reverse=True)
        print("   stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
