#!/usr/bin/env python
"""Fixtures for the phenomenological sliding window: tests/golden/phenom/<case>.npz.

Run in the build container only (needs /root/reference).  The UNMODIFIED reference function
``sliding_window_bposd_phenom_mem`` (decoder/bposd.py:10, sliding_window.py:14-101) is run on detection events sampled by
the oracle, with the oracle's C BP+OSD-0 behind the ldpc shim (fp64 = what ldpc computes in, fp32 = the GPU's other mode).
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

shims.install()
sys.path.insert(0, "/root/reference/src")
import stim  # noqa: E402  (the shim)
from quits.decoder import sliding_window_bposd_phenom_mem  # noqa: E402
from quits.simulation import get_stim_mem_result  # noqa: E402
from tools.make_golden import load  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
BP = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")
# (circuit fixture, W, F, shots, seed, error_rate)
CASES = [
    ("bb72_r6_p1e-3", 5, 3, 384, 31, 0.01),
    ("bb72_r6_p3e-3", 4, 2, 256, 32, 0.03),
    ("bb72_r15_p1e-3", 5, 3, 128, 33, 0.01),
    ("bb144_r10_p1e-3", 5, 3, 192, 34, 0.01),
    ("hgp225_r3_p1e-2", 3, 2, 96, 35, 0.05),
    ("hgp225_r3_p1e-2", 6, 3, 64, 36, 0.05),          # W > rounds + 2: whole-history window (the reference warns)
    ("toric3_zxcol_r3_p1e-3", 3, 1, 512, 37, 0.02),
]


def main():
    os.makedirs(os.path.join(G, "phenom"), exist_ok=True)
    for name, W, F, shots, seed, rate in CASES:
        text, hz, lz = load(name)
        circ = stim.Circuit(text)
        det, obs = get_stim_mem_result(circ, shots, seed=seed)
        preds = {}
        for prec in ("f32", "f64"):
            shims.DEFAULT_PRECISION = prec
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                preds[prec] = sliding_window_bposd_phenom_mem(det, hz, lz, W, F, error_rate=rate, **BP)
        shims.DEFAULT_PRECISION = "f64"
        case = "%s_W%dF%d" % (name, W, F)
        np.savez_compressed(os.path.join(G, "phenom", case + ".npz"), seed=np.int64(seed), shots=np.int64(shots), W=np.int64(W),
                            F=np.int64(F), error_rate=np.float64(rate), det=np.packbits(det, axis=1, bitorder="little"),
                            obs=np.packbits(obs, axis=1, bitorder="little"), D=np.int64(det.shape[1]), K=np.int64(obs.shape[1]),
                            pred_f32=preds["f32"].astype(np.uint8), pred_f64=preds["f64"].astype(np.uint8))
        pl = {k: float(np.mean(np.any((obs.astype(int) - v) % 2, axis=1))) for k, v in preds.items()}
        print("%-28s shots %d  pL(f32) %.4f pL(f64) %.4f  f32!=f64 on %d shots" % (
            case, shots, pl["f32"], pl["f64"], int(np.any(preds["f32"] != preds["f64"], axis=1).sum())))


if __name__ == "__main__":
    main()
