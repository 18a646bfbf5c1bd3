#!/usr/bin/env python
"""Per-source-line view of an `ncu --set full --import-source on` capture: warp instructions, samples, shared-memory
wavefronts and the top stall reasons of the hottest CUDA source lines (ncu's own source page, aggregated rows).

usage: by_line.py <report.ncu-rep> [top_n]      (needs -lineinfo in the build; runs `ncu -i` on this machine)"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    print("== %s" % rep.split("/")[-1])
    print("   %s" % rows[1][1][:150])
    hdr = rows[2]
    isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
    iw = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    iex = hdr.index("L1 Wavefronts Shared Excessive") if "L1 Wavefronts Shared Excessive" in hdr else None
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = [r for r in rows[3:] if len(r) > isamp and r[2] == "-"]          # the per-source-line rows
    num = lambda r, i: 0 if i is None else int(r[i] or 0)   # noqa: E731
    ti, ts, tw = (max(sum(num(r, i) for r in agg), 1) for i in (iinst, isamp, iw))
    print("   warp instructions %d, samples %d, shared-memory wavefronts %d" % (ti, ts, tw))
    print("%6s %7s %7s %7s %9s  %-34s %s" % ("line", "inst%", "samp%", "wavef%", "excess", "top stalls", "source"))
    agg.sort(key=lambda r: -num(r, isamp))
    for r in agg[:top]:
        st = sorted(((hdr[i][6:], num(r, i)) for i in stall), key=lambda kv: -kv[1])[:2]
        print("%6s %7.2f %7.2f %7.2f %9d  %-34s %s" % (
            r[0], 100.0 * num(r, iinst) / ti, 100.0 * num(r, isamp) / ts, 100.0 * num(r, iw) / tw, num(r, iex),
            ", ".join("%s=%d" % kv for kv in st if kv[1]), r[1].strip()[:90]))


if __name__ == "__main__":
    main()
