#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (no GPU needed): compile a .cu to a cubin with -lineinfo, run
nvdisasm --print-line-info, and attribute every instruction to the last '//## File ..., line N' marker.

    python tools/sass_lines.py quits_b200/csrc/bp.cu bp_kernel_compactIdLi512ELi2ELb0 [top]

Complements ncu's per-line view (tools/ncu_lines.py needs a capture): useful to see what a source line costs in every template
variant before spending GPU time.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    src, pattern = os.path.abspath(sys.argv[1]), sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    with tempfile.TemporaryDirectory() as td:
        cubin = os.path.join(td, "k.cubin")
        subprocess.run(["nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--fmad=false", "-cubin",
                        "-o", cubin, src], check=True, cwd=os.path.dirname(src))
        txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], check=True, capture_output=True, text=True).stdout
    lines = open(src).read().split("\n")
    base = os.path.basename(src)
    for fn in re.split(r"\n\s*\.text\.", txt):
        if pattern not in fn[:300]:
            continue
        cur, cnt, total = None, collections.Counter(), 0
        for ln in fn.split("\n"):
            m = re.search(r'//## File "(.*)", line (\d+)', ln)
            if m:
                cur = int(m.group(2)) if m.group(1).endswith(base) else None
                continue
            if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                total += 1
                if cur:
                    cnt[cur] += 1
        print("%s: %d SASS instructions, %d attributed to %s" % (fn.split("\n", 1)[0][:120], total, sum(cnt.values()), base))
        for line, c in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
            print("%5d %5d  %s" % (line, c, lines[line - 1].strip()[:120]))


if __name__ == "__main__":
    main()
