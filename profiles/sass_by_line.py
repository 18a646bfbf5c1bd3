#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS dump by CUDA source line.

usage: sass_by_line.py <src.csv from ncu -i X.ncu-rep --page source --csv> <nvdisasm -g -c output> <kernel substring>
Correlates instruction offsets with the `//## File ... line N` annotations of nvdisasm (needs -lineinfo).
"""
import csv
import re
import sys
from collections import defaultdict


def line_map(sass_path, kernel):
    m, cur, active = {}, None, False
    for ln in open(sass_path):
        if ln.startswith("//---") and ".text." in ln:
            active = kernel in ln
        if not active:
            continue
        mm = re.search(r'//## File "([^"]*)", line (\d+)(.*)', ln)
        if mm:
            # lines of other files (CUDA headers with inlined intrinsics) are reported as "<header>:<line>"
            f = mm.group(1).rsplit("/", 1)[-1]
            cur = int(mm.group(2)) if f.endswith(".cu") else "%s:%s" % (f, mm.group(2))
            continue
        mm = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if mm:
            m[int(mm.group(1), 16)] = cur
    return m


def main():
    src, sass, kernel = sys.argv[1:4]
    lm = line_map(sass, kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = int(rows[2][ia], 16)
    inst, samp = defaultdict(int), defaultdict(int)
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    stalls = defaultdict(lambda: defaultdict(int))
    for r in rows[2:]:
        if len(r) <= isamp or not r[ia].startswith("0x"):
            continue
        line = lm.get(int(r[ia], 16) - base)
        inst[line] += int(r[ii] or 0)
        samp[line] += int(r[isamp] or 0)
        for c in stall_cols:
            if c < len(r) and r[c] not in ("", "0"):
                stalls[line][hdr[c]] += int(r[c])
    tot_i, tot_s = sum(inst.values()), sum(samp.values())
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    print("%6s %12s %6s %8s %6s  top stalls" % ("line", "warp-inst", "%", "samples", "%"))
    for line in sorted(inst, key=lambda k: -samp[k])[:int(sys.argv[4]) if len(sys.argv) > 4 else 40]:
        top = sorted(stalls[line].items(), key=lambda kv: -kv[1])[:3]
        print("%6s %12d %6.2f %8d %6.2f  %s" % (line, inst[line], 100.0 * inst[line] / max(tot_i, 1), samp[line],
                                                100.0 * samp[line] / max(tot_s, 1),
                                                ", ".join("%s=%d" % kv for kv in top)))


if __name__ == "__main__":
    main()
