#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one block per kernel launch with the metrics DESIGN.md / bench.py cite.
usage: ncu_summary.py raw.csv [--update-traffic profiles/r01_e_ncu_traffic.json --capture NAME]"""
import csv
import json
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {}
    for i, h in enumerate(hdr):
        col.setdefault(h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1][0].isupper() else h, i)
        col.setdefault(h, i)
    out = {}
    for r in rows[2:]:
        full = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        name = full.split("<")[0]
        if name in ("k_ed_sign", "k_ed_expand") and "<1>" in full:      # template <bool CT>: the constant-time instantiation
            name += "_ct"
        print(name)
        rec = {}
        for w in WANT:
            i = next((j for j, h in enumerate(hdr) if h == w or h.endswith("." + w)), None)
            if i is None or r[i] == "":
                continue
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:                  # "no data" (e.g. a launch that exits at once)
                continue
            print("   %s = %f %s" % (w, v, units[i]))
            rec[w] = v * SCALE.get(units[i], 1.0) if units[i] in SCALE else v
        if name not in out or rec.get("gpu__time_duration.sum", 0) > out[name].get("gpu__time_duration.sum", 0):
            out[name] = rec                      # several launches of one kernel (tree levels): keep the longest
        print()
    if "--update-traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--update-traffic") + 1]
        cap = sys.argv[sys.argv.index("--capture") + 1] if "--capture" in sys.argv else ""
        j = json.load(open(path))
        for name, rec in out.items():
            rd, wr = rec.get("dram__bytes_read.sum", 0.0), rec.get("dram__bytes_write.sum", 0.0)
            j["kernels"][name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                                  "duration_ms": rec.get("gpu__time_duration.sum"), "grid": int(rec.get("launch__grid_size", 0)),
                                  "registers": int(rec.get("launch__registers_per_thread", 0)),
                                  "fmaheavy_pct": rec.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                                  "alu_pct": rec.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"), "capture": cap}
        json.dump(j, open(path, "w"), indent=1)


main()
