#!/usr/bin/env python
"""BP-LSD fixtures: the UNMODIFIED reference wrapper ``sliding_window_bplsd_circuit_mem`` (decoder/bplsd.py:54-86, loop in
sliding_window.py:104-188) run on the detection events of the committed decode fixtures, with the oracle's C BP + LSD-0
behind the ``ldpc.bplsd_decoder.BpLsdDecoder`` shim.  Build container only (needs /root/reference):

    python tools/make_golden_lsd.py

Writes tests/golden/decode_lsd/<case>.npz = {pred_f32, pred_f64, max_iter}; the inputs are tests/golden/decode/<case>.npz.
The same for the phenomenological twin ``sliding_window_bplsd_phenom_mem`` (decoder/bplsd.py:10-51) on the inputs of
tests/golden/phenom/<case>.npz -> tests/golden/phenom_lsd/<case>.npz.  Each file also carries ``pred_f64_cs1``: the same call with
``lsd_order=1`` (what the reference's own LSD test and doc/05 ask for), over the oracle's per-cluster candidate sweep.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import shims  # noqa: E402

shims.install()
sys.path.insert(0, "/root/reference/src")
import stim  # noqa: E402  (the shim)
from quits.decoder import sliding_window_bplsd_circuit_mem, sliding_window_bplsd_phenom_mem  # noqa: E402

from conftest import case_circuit, circuit_meta, circuit_text, decode_case  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
# (decode fixture, BP iterations): few iterations leave more shots to LSD
CASES = [("qt633_zxcol_r12_p1e-3_W5F3", 10),        # BASELINE config 4: quantum-Tanner 633 code, 12 rounds, BP-LSD inner decoder
         ("bb144_r10_p1e-3_W5F3", 10), ("bb144_r10_p3e-3_W5F3", 4), ("bb72_r6_p3e-3_W5F3", 3), ("hgp225_r3_p1e-2_W3F2", 5),
         ("toric3_zxcol_r3_p1e-3_W3F2", 2)]
PHENOM_CASES = [("bb72_r6_p3e-3_W4F2", 3), ("bb144_r10_p1e-3_W5F3", 4), ("hgp225_r3_p1e-2_W3F2", 3), ("hgp225_r3_p1e-2_W6F3", 3),
                ("toric3_zxcol_r3_p1e-3_W3F1", 2)]


def main():
    os.makedirs(os.path.join(G, "decode_lsd"), exist_ok=True)
    for case, max_iter in CASES:
        g = decode_case(case)
        name = case_circuit(case)
        _, hz, lz = circuit_meta(name)
        circ = stim.Circuit(circuit_text(name))
        preds = {}
        for prec in ("f32", "f64"):
            shims.DEFAULT_PRECISION = prec
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                preds[prec] = sliding_window_bplsd_circuit_mem(g["det"], circ, hz, lz, g["W"], g["F"], max_iter=max_iter, lsd_order=0,
                                                               bp_method="minimum_sum", schedule="parallel", lsd_method="lsd_cs")
        shims.DEFAULT_PRECISION = "f64"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cs1 = sliding_window_bplsd_circuit_mem(g["det"], circ, hz, lz, g["W"], g["F"], max_iter=max_iter, lsd_order=1,
                                                   bp_method="minimum_sum", schedule="parallel", lsd_method="lsd_cs")
        pl = {k: float(np.mean(np.any((g["obs"].astype(int) - v) % 2, axis=1))) for k, v in preds.items()}
        np.savez_compressed(os.path.join(G, "decode_lsd", case + ".npz"), max_iter=np.int64(max_iter),
                            pred_f32=preds["f32"].astype(np.uint8), pred_f64=preds["f64"].astype(np.uint8), pred_f64_cs1=cs1.astype(np.uint8))
        print("%-30s shots %d  pL(f32) %.4f  pL(f64) %.4f  (BP-OSD-0 fixture: %.4f)  lsd_cs 1: pL %.4f, differs from LSD-0 on %d shots" % (
            case, g["shots"], pl["f32"], pl["f64"], float(np.mean(np.any((g["obs"].astype(int) - g["pred_f64"]) % 2, axis=1))),
            float(np.mean(np.any((g["obs"].astype(int) - cs1) % 2, axis=1))), int(np.any(cs1 != preds["f64"], axis=1).sum())))
    os.makedirs(os.path.join(G, "phenom_lsd"), exist_ok=True)
    for case, max_iter in PHENOM_CASES:
        z = np.load(os.path.join(G, "phenom", case + ".npz"))
        D = int(z["D"])
        det = np.unpackbits(z["det"], axis=1, bitorder="little")[:, :D].astype(np.bool_)
        _, hz, lz = circuit_meta(case.rsplit("_W", 1)[0])
        preds = {}
        for prec in ("f32", "f64"):
            shims.DEFAULT_PRECISION = prec
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                preds[prec] = sliding_window_bplsd_phenom_mem(det, hz, lz, int(z["W"]), int(z["F"]), float(z["error_rate"]), max_iter=max_iter,
                                                              lsd_order=0, bp_method="minimum_sum", schedule="parallel", lsd_method="lsd_cs")
        shims.DEFAULT_PRECISION = "f64"
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cs1 = sliding_window_bplsd_phenom_mem(det, hz, lz, int(z["W"]), int(z["F"]), float(z["error_rate"]), max_iter=max_iter,
                                                  lsd_order=1, bp_method="minimum_sum", schedule="parallel", lsd_method="lsd_cs")
        np.savez_compressed(os.path.join(G, "phenom_lsd", case + ".npz"), max_iter=np.int64(max_iter),
                            pred_f32=preds["f32"].astype(np.uint8), pred_f64=preds["f64"].astype(np.uint8), pred_f64_cs1=cs1.astype(np.uint8))
        print("phenom %-30s shots %d  differs from the BP-OSD-0 fixture on %d shots" % (
            case, det.shape[0], int(np.any(preds["f64"] != z["pred_f64"], axis=1).sum())))


if __name__ == "__main__":
    main()
