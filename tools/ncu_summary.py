#!/usr/bin/env python3
"""Condense an `ncu --page raw --csv` export into the per-kernel summary kept under profiles/.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/rN_ncu_full_summary.csv [profiles/traffic.json]

Keeps the metrics the roofline discussion in DESIGN.md uses (duration, DRAM bytes, instruction and
pipe utilisation, shared-memory wavefronts and conflicts, stall ratios, launch geometry) and writes
traffic.json = {kernel base name: dram read + write bytes per launch}, which bench.py reports as
roofline.traffic.
"""
import csv
import json
import re
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
]


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [h for h in KEEP if h in hdr] + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled")]
    traffic = {}
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel", "metric", "unit", "value"])
        for r in data:
            name = r[hdr.index("Kernel Name")]
            base = re.sub(r"^void\s+", "", name)
            base = re.sub(r"[<(].*$", "", base).split("::")[-1]
            for c in cols:
                w.writerow([name, c, units[hdr.index(c)], r[hdr.index(c)]])
            def val(metric):
                i = hdr.index(metric)
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)
                return float(r[i]) * scale
            traffic[base] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    if len(sys.argv) > 3:
        with open(sys.argv[3], "w") as fh:
            json.dump(traffic, fh, indent=1)
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
