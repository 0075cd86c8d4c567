#!/usr/bin/env python
"""Turns `ncu --page raw --csv` exports of one EM step (gpurun_out/<dir>/<wl>_step_raw.csv) into the committed
profiles/<round>_<wl>_step_ncu_full_summary.txt and profiles/<round>_traffic.json.

    python tools/ncu_to_profiles.py gpurun_out/r02g r02 c2:1000000 c3s:500000
(the number after the colon is the rows per launch = the chunk size of that capture)."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("time_us", "gpu__time_duration.sum"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"), ("smem_dyn_KB", "launch__shared_mem_per_block_dynamic"),
    ("occ_warps_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("fp64_pipe_pct", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
    ("dmma_pipe_pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), ("inst", "smsp__inst_executed.sum"),
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio$")
SCALE = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "byte": 1e-6, "kbyte": 1e-3, "mbyte": 1, "gbyte": 1e3}


def num(v, unit):
    try:
        return float(v.replace(",", "")) * SCALE.get(unit.lower(), 1)
    except ValueError:
        return float("nan")


def main(directory, rnd, specs):
    traffic = {}
    for spec in specs:
        wl, rows_per_launch = spec.split(":")
        path = os.path.join(directory, f"{wl}_step_raw.csv")
        if not os.path.exists(path):
            path = os.path.join(directory, f"{wl}_full_raw.csv")
        rows = list(csv.reader(open(path)))
        hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        hdr, units = rows[hi], rows[hi + 1]
        idx = {h: i for i, h in enumerate(hdr)}
        lines, fam = [], {}
        for r in rows[hi + 2:]:
            name = r[idx["Kernel Name"]]
            lines.append("== " + name[:110])
            for label, key in KEYS:
                if key in idx:
                    lines.append(f"   {label:18s} {num(r[idx[key]], units[idx[key]]):14.3f}")
            st = []
            for h, i in idx.items():
                m = STALL.match(h)
                if m and r[i] not in ("", "n/a"):
                    try:
                        st.append((float(r[i].replace(",", "")), m.group(1)))
                    except ValueError:
                        pass
            lines.append("   stalls: " + ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))
            short = re.sub(r"^void ", "", name).split("(")[0]
            fam.setdefault(short, []).append((num(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) * 1e6,
                                              num(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) * 1e6,
                                              num(r[idx["gpu__time_duration.sum"]], units[idx["gpu__time_duration.sum"]])))
        with open(os.path.join(ROOT, "profiles", f"{rnd}_{wl}_step_ncu_full_summary.txt"), "w") as f:
            f.write(f"# ncu --set full --clock-control none, consecutive launches of one EM step of bench.py --workload {wl} "
                    f"({rows_per_launch} rows per launch)\n" + "\n".join(lines) + "\n")
        traffic[wl] = {"rows_per_launch": int(rows_per_launch),
                       "source": f"ncu --set full --clock-control none on bench.py --workload {wl}; summary in "
                                 f"profiles/{rnd}_{wl}_step_ncu_full_summary.txt",
                       "kernels": {n: {"launches": len(v), "dram_bytes_per_launch": sum(a + b for a, b, _ in v) / len(v),
                                       "dram_read_per_launch": sum(a for a, _, _ in v) / len(v),
                                       "dram_write_per_launch": sum(b for _, b, _ in v) / len(v),
                                       "time_us_under_ncu": sum(t for _, _, t in v) / len(v)} for n, v in fam.items()}}
        for n, v in traffic[wl]["kernels"].items():
            print(wl, n[:48], v["launches"], f"{v['dram_bytes_per_launch'] / 1e6:.1f} MB", f"{v['time_us_under_ncu']:.1f} us")
    with open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3:])
