#!/usr/bin/env python
"""Where the end-to-end time of the drop-in calls goes (GPU box): sample / decode / caller's numpy reduction, host wall clock."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import quits_b200 as qb  # noqa: E402
from conftest import circuit_meta, circuit_text  # noqa: E402

name = "bb144_r10_p1e-3"
_, hz, lz = circuit_meta(name)
c = qb.Circuit(circuit_text(name))
KW = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
for i in range(4):
    t0 = time.perf_counter()
    det, obs = qb.get_stim_mem_result(c, S, seed=100 + i)
    t1 = time.perf_counter()
    pred = qb.sliding_window_bposd_circuit_mem(det, c, hz, lz, 5, 3, **KW)
    t2 = time.perf_counter()
    nerr = int(np.any((obs - pred) % 2, axis=1).sum())
    t3 = time.perf_counter()
    print("step %d: sample %.1f ms  decode %.1f ms  reduce %.1f ms  total %.1f ms  (%d errors)" % (
        i, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t3 - t0), nerr))
