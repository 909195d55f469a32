#!/usr/bin/env python
"""The reference arm proper: mkangquantum/quits UNMODIFIED (installed under baseline/_ref with
`pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`) driving the real `stim` and `ldpc`
wheels on the host cores, for the workload bench.py measures.

    python baseline/run_reference.py --probe                     # exit 0 and print {"available": true} if stim+ldpc import
    python baseline/run_reference.py --workload bb144_r10_p1e-3 --shots 2000 [--procs N] [--doc-setting]
    python baseline/run_reference.py --pin tests/golden/pinned   # regenerate the external-pin fixtures from the real wheels

The reference call sites driven here, untouched: src/quits/simulation.py:22-27 (stim detector sampler),
src/quits/decoder/sliding_window.py:146-153,171,182 (one ldpc BpOsdDecoder per window, one decode per shot and window),
dependencies pyproject.toml:29-36 (stim>=1.13.0, ldpc>=2.1.2 -- neither is vendored, pinned or present in this image's
wheelhouse, so on this image the script reports "unavailable" and bench.py falls back to the oracle's C port, saying so).

One Python process per host core (the reference is single-threaded per call), each with its own seed, N shots split evenly;
the figure is shots / wall-clock of the slowest worker, imports and decoder construction excluded (as in bench.py's own arm).
"""
from __future__ import annotations

import argparse
import gzip
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def probe():
    """(ok, why): are the real stim / ldpc wheels and the installed reference importable?"""
    if os.path.isdir(REF) and REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        import stim          # noqa: F401
    except Exception as e:   # pragma: no cover - depends on the image
        return False, "stim not importable (%s: %s); not in /opt/wheelhouse, no network" % (type(e).__name__, e)
    try:
        import ldpc          # noqa: F401
        from ldpc.bposd_decoder import BpOsdDecoder   # noqa: F401
    except Exception as e:   # pragma: no cover
        return False, "ldpc not importable (%s: %s); not in /opt/wheelhouse, no network" % (type(e).__name__, e)
    try:
        import quits         # noqa: F401
    except Exception as e:   # pragma: no cover
        return False, "reference package not importable from baseline/_ref (%s: %s)" % (type(e).__name__, e)
    return True, ""


def load_workload(name):
    import numpy as np
    g = os.path.join(ROOT, "tests", "golden", "circuits")
    p = os.path.join(g, name + ".stim")
    if os.path.exists(p):
        text = open(p).read()
    else:
        with gzip.open(p + ".gz", "rb") as f:
            text = f.read().decode()
    meta = json.load(open(os.path.join(g, name + ".json")))
    hz = np.zeros(meta["hz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["hz_rows"]):
        hz[i, r] = 1
    lz = np.zeros(meta["lz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["lz_rows"]):
        lz[i, r] = 1
    return text, hz, lz


def _worker(args):
    name, shots, seed, W, F, kw = args
    import numpy as np
    import stim
    from quits.decoder import sliding_window_bposd_circuit_mem
    from quits.simulation import get_stim_mem_result
    text, hz, lz = load_workload(name)
    circuit = stim.Circuit(text)
    t0 = time.perf_counter()
    det, obs = get_stim_mem_result(circuit, shots, seed=seed)
    pred = sliding_window_bposd_circuit_mem(det, circuit, hz, lz, W, F, **kw)
    fails = int(np.any((obs - pred) % 2, axis=1).sum())
    return time.perf_counter() - t0, fails


def run(name, shots, procs, W, F, kw, seed=20260101):
    procs = max(1, procs)
    per = max(1, shots // procs)
    jobs = [(name, per, seed + i, W, F, kw) for i in range(procs)]
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_worker, [(name, 2, seed + 999, W, F, kw)] * procs)        # imports + one warm call per worker, untimed
        t0 = time.perf_counter()
        res = pool.map(_worker, jobs)
        wall = time.perf_counter() - t0
    return {"shots": per * procs, "wall_s": wall, "slowest_worker_s": max(r[0] for r in res), "shots_per_s": per * procs / wall,
            "procs": procs, "logical_errors": int(sum(r[1] for r in res))}


def pin(outdir):
    """External pin: small fixtures produced by the real wheels (DEM digest, priors, detection events, predictions for osd_0 /
    osd_cs 1 / lsd) that tests/test_pinned.py compares the oracle and the GPU path with when they exist."""
    import hashlib
    import numpy as np
    import stim
    from quits.decoder import sliding_window_bposd_circuit_mem, sliding_window_bplsd_circuit_mem
    from quits.decoder.base import detector_error_model_to_matrix
    os.makedirs(outdir, exist_ok=True)
    for name, W, F in (("bb72_r6_p3e-3", 5, 3), ("bb144_r10_p1e-3", 5, 3)):
        text, hz, lz = load_workload(name)
        c = stim.Circuit(text)
        dem = c.detector_error_model(decompose_errors=False)
        H, L, pri = detector_error_model_to_matrix(dem)
        det, obs = c.compile_detector_sampler(seed=7).sample(shots=64, separate_observables=True)
        out = {"stim": stim.__version__, "H_sha": hashlib.sha256(H.tocsc().indices.tobytes()).hexdigest(),
               "priors_sha": hashlib.sha256(np.asarray(pri, dtype=np.float64).tobytes()).hexdigest(), "ncols": int(H.shape[1]),
               "det": np.packbits(det, axis=1).tolist()}
        for tag, kw in (("osd0_ms_par", dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")),
                        ("oscs1_ps_ser", dict(max_iter=10, osd_order=1, bp_method="product_sum", schedule="serial", osd_method="osd_cs"))):
            out[tag] = sliding_window_bposd_circuit_mem(det, c, hz, lz, W, F, **kw).astype(int).tolist()
        out["lsd1"] = sliding_window_bplsd_circuit_mem(det, c, hz, lz, W, F, max_iter=10, lsd_order=1, bp_method="minimum_sum",
                                                       schedule="parallel", lsd_method="lsd_cs").astype(int).tolist()
        json.dump(out, open(os.path.join(outdir, name + ".json"), "w"))
        print("pinned", name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--probe", action="store_true")
    ap.add_argument("--pin", default=None)
    ap.add_argument("--workload", default="bb144_r10_p1e-3")
    ap.add_argument("--shots", type=int, default=512)
    ap.add_argument("--procs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--doc-setting", action="store_true", help="product_sum / serial / osd_cs order 1 (the notebooks) instead of the headline")
    args = ap.parse_args()
    ok, why = probe()
    if args.probe or not ok:
        print(json.dumps({"available": ok, "why": why}))
        return 0
    if args.pin:
        pin(args.pin)
        return 0
    kw = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")
    if args.doc_setting:
        kw = dict(max_iter=10, osd_order=1, bp_method="product_sum", schedule="serial", osd_method="osd_cs")
    r = run(args.workload, args.shots, args.procs, 5, 3, kw)
    r["available"] = True
    print(json.dumps(r))
    return 0


if __name__ == "__main__":
    sys.exit(main())
