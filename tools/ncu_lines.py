#!/usr/bin/env python
"""Per-source-line summary of an ncu report (compiled with -lineinfo, captured with --import-source on).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n]

Prints, for the hottest CUDA source lines of the first matching kernel: warp-instructions executed, stall samples,
shared-memory wavefronts (total / ideal).  Runs on the CPU box: it only reads the report.
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    lines = []
    for r in rows:
        if len(r) > 3 and r[0] == "Line No":
            if hdr is not None:
                break           # second kernel instance: stop
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        lines.append(r)
    col = {h: i for i, h in enumerate(hdr)}
    ci, cs = col["Instructions Executed"], col["# Samples"]
    cw, cwi = col.get("L1 Wavefronts Shared"), col.get("L1 Wavefronts Shared Ideal")
    tot_i = sum(float(r[ci]) for r in lines)
    tot_s = sum(float(r[cs]) for r in lines)
    tot_w = sum(float(r[cw]) for r in lines)
    print("total: %.3g warp-instr, %d samples, %.3g smem wavefronts" % (tot_i, tot_s, tot_w))
    lines.sort(key=lambda r: -float(r[cs]))
    print("%5s %7s %7s %9s %9s  %s" % ("line", "inst%", "samp%", "smemWF%", "WF/ideal", "source"))
    for r in lines[:top]:
        w, wi = float(r[cw]), float(r[cwi])
        print("%5s %7.2f %7.2f %9.2f %9.2f  %s" % (r[0], 100 * float(r[ci]) / max(tot_i, 1), 100 * float(r[cs]) / max(tot_s, 1),
                                                    100 * w / max(tot_w, 1), w / wi if wi else 0.0, r[1].strip()[:110]))


if __name__ == "__main__":
    main()
