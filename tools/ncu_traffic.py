#!/usr/bin/env python3
"""Regenerates profiles/ncu_traffic.json from ONE `ncu --set full` capture of the benchmark workload (developer tool).

Usage: ncu_traffic.py <rep.ncu-rep> <segments in the captured launch> [kernel-regex]
The capture command (GPU box, after the same command exited 0 without ncu):
  ncu --set full --clock-control none --import-source on -k regex:k_render_(fused|pool) --launch-skip 1 --launch-count 1 \
      -o gpurun_out/<name> python tools/prof_run.py cornell-box-scene.json 1024
bench.py reads the per-segment DRAM bytes, the issue-slot utilisation and the lanes active per instruction from the
JSON and names this capture as their source."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, segs = sys.argv[1], float(sys.argv[2])
    kre = sys.argv[3] if len(sys.argv) > 3 else "k_render"
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        data = json.load(open(path))
    except (OSError, ValueError):
        data = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if kre not in name and not __import__("re").search(kre, name):
            continue
        base = name.split("<")[0].replace("void ", "").strip()

        def f(key):
            return float(r[idx[key]].replace(",", ""))
        unit_r, unit_w = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        dram = f("dram__bytes_read.sum") * mult[unit_r] + f("dram__bytes_write.sum") * mult[unit_w]
        data[base] = {
            "source": "profiles/" + os.path.basename(rep).replace(".ncu-rep", "") + "_metrics.csv",
            "kernel": name, "segments_in_capture": segs, "gpu_time_ms": f("gpu__time_duration.sum") *
            {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[rows[1][idx["gpu__time_duration.sum"]]],
            "dram_bytes": dram, "dram_bytes_per_segment": dram / segs,
            "issue_slot_utilisation_pct": f("sm__inst_issued.avg.pct_of_peak_sustained_active"),
            "active_threads_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "registers_per_thread": f("launch__registers_per_thread"),
            "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_rate_pct": f("l1tex__t_sector_hit_rate.pct"), "l2_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
        }
        # keep the raw metric row next to the JSON so every number can be traced
        with open(os.path.join(ROOT, data[base]["source"]), "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow(hdr)
            w.writerow(rows[1])
            w.writerow(r)
        print(base, json.dumps(data[base], indent=1))
    json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
