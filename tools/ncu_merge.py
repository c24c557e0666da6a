#!/usr/bin/env python3
"""Condense the captures of tools/ncu_capture.sh into profiles/ (run in the build container, where ncu reads reports).

    python tools/ncu_merge.py r2g r2        # gpurun_out/r2g_{build,apply}_{r1,r8seg}.ncu-rep -> profiles/r2_ncu_full_summary.csv,
                                            # profiles/traffic.json, profiles/r2_launches.csv

One row per (capture, kernel, metric): the metrics the roofline discussion in DESIGN.md uses (tools/ncu_summary.py
keeps the list).  traffic.json = DRAM bytes per base of every captured shape; bench.py scales it to the batch it
times and reports it as roofline.traffic."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import KEEP  # noqa: E402

BENCH = "python bench.py --steps 2 --warmup 3 --no-configs --no-fastq --no-cpu --no-e2e"
SHAPES = {"r1": ("L150_R1_read-order", 10_000_000, BENCH),
          "r8seg": ("L150_R8_segmented", 25_000_000, BENCH + " --read-groups 8 --reads 25000000")}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    tag, rnd = sys.argv[1], sys.argv[2]
    out_csv = os.path.join(ROOT, "profiles", "%s_ncu_full_summary.csv" % rnd)
    traffic = {}
    with open(out_csv, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["capture", "kernel", "metric", "unit", "value"])
        for shape, (key, reads, cmd) in SHAPES.items():
            entry = {"capture": "profiles/%s_ncu_full_summary.csv (ncu --set full --clock-control none of `%s`)" % (rnd, cmd),
                     "reads": reads}
            for kern in ("build", "apply"):
                rep = os.path.join(ROOT, "gpurun_out", "%s_%s_%s.ncu-rep" % (tag, kern, shape))
                hdr, units, data = raw_rows(rep)
                cols = [h for h in KEEP if h in hdr] + [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled")]
                r = data[0]
                name = r[hdr.index("Kernel Name")]
                for c in cols:
                    w.writerow(["%s_%s_%s" % (tag, kern, shape), name, c, units[hdr.index(c)], r[hdr.index(c)]])

                def val(metric):
                    i = hdr.index(metric)
                    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(units[i], 1.0)
                    return float(r[i]) * scale
                dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
                dur = hdr.index("gpu__time_duration.sum")
                entry["%s_smem_kernel" % kern] = {"bytes_per_base": dram / (reads * 150), "dram_bytes": dram,
                                                  "duration_us": float(r[dur]) * {"ms": 1e3, "us": 1.0, "ns": 1e-3}.get(units[dur], 1.0),
                                                  "algorithmic_bytes_per_base": 3}
            traffic[key] = entry
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fh:
        json.dump(traffic, fh, indent=1)
    # launch list of the bench command: id, kernel, grid, block, duration; shares per kernel on stdout
    raw = os.path.join(ROOT, "gpurun_out", "%s_launches_raw.csv" % tag)
    if os.path.exists(raw):
        rows = list(csv.reader(l for l in open(raw) if l.startswith('"')))
        hdr = rows[0]
        col = {h: hdr.index(h) for h in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Unit", "Metric Value")}
        tot = {}
        with open(os.path.join(ROOT, "profiles", "%s_launches.csv" % rnd), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(["id", "kernel", "grid", "block", "gpu__time_duration_us"])
            for r in rows[1:]:
                us = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[col["Metric Unit"]], 1.0)
                name = re.sub(r"\(.*$", "", re.sub(r"^void\s+", "", r[col["Kernel Name"]]))
                w.writerow([r[col["ID"]], name, r[col["Grid Size"]], r[col["Block Size"]], "%.3f" % us])
                base = re.sub(r"<.*$", "", name)
                tot[base] = tot.get(base, 0.0) + us
        total = sum(tot.values())
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            print("%-60s %12.1f us  %5.1f %%" % (k, v, 100 * v / total))
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
