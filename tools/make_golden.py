#!/usr/bin/env python
"""Generate the parity fixtures under tests/golden/ by running the UNMODIFIED reference on top of the oracle.

Run in the build container only (needs /root/reference):

    python tools/make_golden.py

For every frozen circuit text (tests/golden/circuits/*.stim, made by tools/make_circuits.py) and window setting:
  * windows/<case>.json   digest of what the reference's own ``spacetime()`` (decoder/base.py:134-190) returns when
                          fed the oracle's DEM through the stim shim: shapes, nnz, sha256 of indptr / indices / priors
  * decode/<case>.npz     seed, detection events and observable flips of the oracle sampler (bit-packed), and the
                          prediction of the reference's ``sliding_window_bposd_circuit_mem`` (decoder/bposd.py:54,
                          sliding_window.py:104-188) with the oracle's C BP+OSD-0 behind the ldpc shim, in fp32
                          (what the GPU is held to, bit for bit) and fp64 (what ldpc computes in)
  * dem/<circuit>.json    digest of the stim-ordered DEM (error count, sha256 of probabilities and targets)
The GPU box never runs this; it reads the committed files.
"""
import hashlib
import json
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
from quits.decoder import sliding_window_bposd_circuit_mem, spacetime  # noqa: E402
from quits.simulation import get_stim_mem_result  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
BP = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")

# (circuit fixture, W, F, shots, seed)
CASES = [
    ("bb72_r6_p1e-3", 5, 3, 512, 11),
    ("bb72_r6_p3e-3", 5, 3, 384, 12),
    ("bb72_r15_p1e-3", 5, 3, 192, 13),
    ("bb72_r3_p1e-3_X", 3, 2, 256, 14),
    ("bb144_r10_p1e-3", 5, 3, 512, 20260101),
    ("bb144_r10_p3e-3", 5, 3, 192, 15),
    ("bb144_r10_p3e-4", 5, 3, 512, 16),
    ("bb144_r10_p1e-3", 10, 5, 128, 17),
    ("hgp225_r3_p1e-2", 3, 2, 128, 18),
    ("hgp225_r3_p1e-2", 5, 3, 96, 19),          # W > rounds: whole-history window (the reference warns)
    ("toric3_zxcol_r3_p1e-3", 3, 2, 1024, 21),
    ("qt633_zxcol_r12_p1e-3", 5, 3, 128, 22),    # BASELINE config 4 (circuit from tools/make_circuits_more.py), BP-OSD inner decoder
    ("hgp225_r15_p1e-3", 5, 3, 48, 23),          # the notebooks' HGP run: 540 x 6480 windows
    ("lcs_card_r6_p1e-3", 5, 3, 256, 24),        # the other code families / strategies of the reference's tests
    ("bpc90_nsmerge_r6_p1e-3", 5, 3, 192, 25),
    ("qlp544_card_r6_p5e-4", 5, 3, 48, 26),      # windows of 1200 checks: generic BP kernel + slab OSD
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load(name):
    path = os.path.join(G, "circuits", name + ".stim")
    if os.path.exists(path):
        with open(path) as f:
            text = f.read()
    else:                                              # the larger texts are committed compressed
        import gzip
        with gzip.open(path + ".gz", "rb") as f:
            text = f.read().decode()
    with open(os.path.join(G, "circuits", name + ".json")) as f:
        meta = json.load(f)
    hz = np.zeros(meta["hz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["hz_rows"]):
        hz[i, r] = 1
    lz = np.zeros(meta["lz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["lz_rows"]):
        lz[i, r] = 1
    return text, hz, lz


def n_cor_rounds(D, m, W, F):
    rounds = D // m - 2
    if 2 + rounds - W >= 0:
        n = (2 + rounds - W) // F
        if (2 + rounds - W) % F:
            n += 1
        return n
    return 0


def main():
    for sub in ("windows", "decode", "dem"):
        os.makedirs(os.path.join(G, sub), exist_ok=True)
    done_dem = set()
    only = sys.argv[1:]                         # optional: substrings of the case names to (re)generate
    for name, W, F, shots, seed in CASES:
        if only and not any(o in name for o in only):
            continue
        text, hz, lz = load(name)
        circ = stim.Circuit(text)
        m = hz.shape[0]
        case = "%s_W%dF%d" % (name, W, F)
        if name not in done_dem:
            done_dem.add(name)
            d = circ.detector_error_model(decompose_errors=False)._dem
            flat_d = np.array([x for e in d.dets for x in e], dtype=np.int32)
            flat_o = np.array([x for e in d.obs for x in e], dtype=np.int32)
            with open(os.path.join(G, "dem", name + ".json"), "w") as f:
                json.dump({"n_errors": len(d.probs), "n_det": d.n_det, "n_obs": d.n_obs, "sum_probs": float(np.sum(d.probs)),
                           "probs_sha256": sha(np.array(d.probs, dtype=np.float64)),
                           "det_len_sha256": sha(np.array([len(e) for e in d.dets], dtype=np.int32)),
                           "det_idx_sha256": sha(flat_d), "obs_len_sha256": sha(np.array([len(e) for e in d.obs], dtype=np.int32)),
                           "obs_idx_sha256": sha(flat_o)}, f, indent=1)
        # ---- the reference's own window slicing
        ncr = n_cor_rounds(circ.num_detectors, m, W, F)
        checks, observables, priors, updates = spacetime(circ, hz, W, F, ncr)
        wins = []
        for k in range(len(checks)):
            H = checks[k].tocsc(); H.sort_indices()
            L = observables[k].tocsc(); L.sort_indices()
            ent = {"H_shape": list(H.shape), "H_nnz": int(H.nnz), "H_indptr": sha(H.indptr.astype(np.int64)),
                   "H_indices": sha(H.indices.astype(np.int32)), "L_shape": list(L.shape), "L_nnz": int(L.nnz),
                   "L_indptr": sha(L.indptr.astype(np.int64)), "L_indices": sha(L.indices.astype(np.int32)),
                   "priors": sha(np.asarray(priors[k], dtype=np.float64)), "priors_sum": float(np.sum(priors[k]))}
            if k < len(updates):
                U = updates[k].tocsc(); U.sort_indices()
                ent.update({"U_shape": list(U.shape), "U_nnz": int(U.nnz), "U_indptr": sha(U.indptr.astype(np.int64)),
                            "U_indices": sha(U.indices.astype(np.int32))})
            wins.append(ent)
        with open(os.path.join(G, "windows", case + ".json"), "w") as f:
            json.dump({"circuit": name, "W": W, "F": F, "m": m, "num_cor_rounds": ncr, "windows": wins}, f, indent=1)
        # ---- sample with the oracle, decode with the reference loop over the oracle's BP+OSD-0
        det, obs = get_stim_mem_result(circ, shots, seed=seed)
        preds = {}
        for prec in ("f32", "f64"):
            shims.DEFAULT_PRECISION = prec
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                preds[prec] = sliding_window_bposd_circuit_mem(det, circ, hz, lz, W, F, **BP)
        shims.DEFAULT_PRECISION = "f64"
        pl = {k: float(np.mean(np.any((obs.astype(int) - v) % 2, axis=1))) for k, v in preds.items()}
        np.savez_compressed(os.path.join(G, "decode", case + ".npz"), seed=np.int64(seed), shots=np.int64(shots), W=np.int64(W), F=np.int64(F),
                            m=np.int64(m), det=np.packbits(det, axis=1, bitorder="little"), obs=np.packbits(obs, axis=1, bitorder="little"),
                            D=np.int64(det.shape[1]), K=np.int64(obs.shape[1]),
                            pred_f32=preds["f32"].astype(np.uint8), pred_f64=preds["f64"].astype(np.uint8))
        print("%-28s windows %s  shots %d  pL(f32) %.4f pL(f64) %.4f  f32!=f64 on %d shots" % (
            case, [tuple(w["H_shape"]) for w in wins], shots, pl["f32"], pl["f64"],
            int(np.any(preds["f32"] != preds["f64"], axis=1).sum())))


if __name__ == "__main__":
    main()
