#!/usr/bin/env python
"""profiles/traffic.json entry from one `ncu --set full` capture.

    ncu -i gpurun_out/prof_c1.ncu-rep --page raw --csv > /tmp/raw.csv
    python scripts/make_traffic.py hf_run_c1_kernel /tmp/raw.csv 200000000 "profiles/r02a_... (what one launch is)" hf_frm_c1.cuh hf_frm.cuh hf_philox.cuh

Each entry is stamped with the sha256 of the source files its kernel is compiled from (`sources`): bench.py drops
`roofline.traffic` and `issue` when the stamp no longer matches what is built, instead of reporting counters of
another kernel version."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hyperfluorescence_b200.build import source_hashes  # noqa: E402


def main():
    key, raw, events, source = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    files = sys.argv[5:]
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
        d[h] = x * scale
    rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
    entry = {
        "kernel": d and rows[2][hdr.index("Kernel Name")], "source": source, "events_per_launch": events,
        "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_launch": rd + wr, "dram_bytes_per_event": (rd + wr) / events,
        "duration_ms": d["gpu__time_duration.sum"],
        "warp_inst_per_event": d["smsp__inst_executed.sum"] / events,
        "issue_slots_busy_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "fp64_pipe_pct": d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "alu_pipe_pct": d.get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "l2_hit_pct": d.get("lts__t_sector_hit_rate.pct"), "l1_hit_pct": d.get("l1tex__t_sector_hit_rate.pct"),
        "achieved_occupancy_pct": d.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": d.get("launch__registers_per_thread"),
        "sources": source_hashes(files),
    }
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        allj = json.load(open(path))
    except Exception:
        allj = {}
    allj[key] = entry
    json.dump(allj, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
