#!/usr/bin/env python
"""Instruction share per CUDA source line (in file order) of the first kernel matching KERNEL_REGEX in an ncu report.

    python tools/ncu_inst.py REPORT.ncu-rep KERNEL_REGEX [min_pct]
"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, lines = None, []
    for r in rows:
        if len(r) > 3 and r[0] == "Line No":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        lines.append(r)
    col = {h: i for i, h in enumerate(hdr)}
    ci, cs = col["Instructions Executed"], col["# Samples"]
    tot = sum(float(r[ci]) for r in lines)
    tots = sum(float(r[cs]) for r in lines)
    lines.sort(key=lambda r: int(r[0]))
    print("total %.4g warp-instructions, %d samples" % (tot, tots))
    for r in lines:
        f = 100 * float(r[ci]) / tot
        if f >= minp:
            print("%4s inst %6.2f%%  samp %6.2f%%  %s" % (r[0], f, 100 * float(r[cs]) / max(tots, 1), r[1].strip()[:110]))


if __name__ == "__main__":
    main()
