#!/usr/bin/env python
"""Per-opcode warp-stall summary of one kernel from `ncu -i rep --page source --csv --print-source sass --kernel-name ...`:
which instructions the sampled warps were waiting on, and why.  usage: ncu_stalls.py source.csv [> profiles/rNN_..._stalls.txt]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    samp, execd, stall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)

    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[idx["Source"]].strip())
        op = ".".join((m.group(2) if m else r[idx["Source"]][:10]).split(".")[:3])
        samp[op] += num(r[idx["# Samples"]])
        execd[op] += num(r[idx["Instructions Executed"]])
        for c in stall_cols:
            v = num(r[idx[c]])
            if v:
                stall[op][c.replace("stall_", "")] += v
    tot, tex = sum(samp.values()), sum(execd.values())
    print(rows[0][1] if len(rows[0]) > 1 else "")
    print("warp-state samples: %d; instructions executed (as the page counts them): %d" % (tot, tex))
    print("%-22s %8s %7s %7s  %s" % ("opcode", "samples", "share", "of exec", "top stall reasons (samples)"))
    for op, s in samp.most_common(14):
        top = ", ".join("%s %d" % kv for kv in stall[op].most_common(4))
        print("%-22s %8d %6.1f%% %6.1f%%  %s" % (op, s, 100.0 * s / max(tot, 1), 100.0 * execd[op] / max(tex, 1), top))
    allst = collections.Counter()
    for op in stall:
        allst.update(stall[op])
    print("all opcodes: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot, 1)) for k, v in allst.most_common(8)))


main()
