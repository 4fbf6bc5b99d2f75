#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed) into profiles/: the handful of counters
DESIGN.md and bench.py quote, per captured kernel launch.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_ncu.txt [--json profiles/x.json --natom 2000 --note "..."]
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import subprocess

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]  # fmt: skip


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--json")
    ap.add_argument("--natom", type=int)
    ap.add_argument("--note", default="")
    ns = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", ns.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"ncu --set full report {ns.report}", ns.note, "Profiled run (serialised, cold caches): use ratios, not times.", ""]
    first = None
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        if first is None:
            first = (d, u)
        lines.append(d.get("Kernel Name", "?")[:150])
        for k in KEYS:
            if k in d and d[k] not in ("", "n/a"):
                lines.append(f"  {k:85s} {d[k]:>18s} {u[k]}")
        stalls = sorted(((float(d[k]), k) for k in d if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and d[k]),
                        reverse=True)[:6]  # fmt: skip
        for v, k in stalls:
            lines.append(f"  {k:85s} {v:18.3f}")
        lines.append("")
    open(ns.out, "w").write("\n".join(lines))
    print("\n".join(lines[:60]))
    if ns.json and first:
        d, u = first
        out = {"kernel": d.get("Kernel Name"), "natom": ns.natom, "source": f"{ns.out} ({ns.note})",
               "dram_bytes_read": to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]),
               "dram_bytes_write": to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]),
               "fp64_pipe_pct": float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]),
               "registers_per_thread": int(d["launch__registers_per_thread"])}  # fmt: skip
        json.dump(out, open(ns.json, "w"), indent=1)


if __name__ == "__main__":
    main()
