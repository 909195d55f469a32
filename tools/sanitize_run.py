#!/usr/bin/env python
"""Small pass over every kernel family, meant to be run under compute-sanitizer on the GPU box:

    compute-sanitizer --tool memcheck python tools/sanitize_run.py
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quits_b200 as qb  # noqa: E402
from bench import load_workload  # noqa: E402


def circuit(name):
    with open(os.path.join(ROOT, "tests", "golden", "circuits", name + ".stim")) as f:
        return qb.Circuit(f.read())


def main():
    warnings.simplefilter("ignore")
    text, hz, lz = load_workload()
    c = qb.Circuit(text)
    scale = float(os.environ.get("QB_SANITIZE_SCALE", "1"))           # racecheck is slow: QB_SANITIZE_SCALE=0.2
    n = lambda x: max(8, int(x * scale))
    det, obs = qb.get_stim_mem_result(c, n(300), seed=3)
    base = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")
    for kw in (base, dict(base, bp_method="product_sum"), dict(base, osd_method="osd_cs", osd_order=2),
               dict(base, osd_method="osd_e", osd_order=3), dict(base, schedule="serial", max_iter=3),
               dict(base, schedule="serial", bp_method="product_sum", max_iter=2)):
        for prec in ("f64", "f32"):
            dec = qb.SlidingWindowDecoder(c, hz.shape[0], 5, 3, precision=prec, **kw)
            pred = dec.decode(det)
            print(kw["bp_method"], kw["schedule"], kw["osd_method"], kw["osd_order"], prec, "logical errors",
                  int(np.any((obs - pred) % 2, axis=1).sum()), flush=True)
    for prec in ("f64", "f32"):                          # BP-LSD (lsd_kernel), few BP iterations so that LSD has work
        dec = qb.SlidingWindowDecoder(c, hz.shape[0], 5, 3, precision=prec, **dict(base, osd_method="lsd_0", max_iter=3))
        print("lsd_0", prec, "logical errors", int(np.any((obs - dec.decode(det)) % 2, axis=1).sum()), flush=True)
        for m_, o_ in (("lsd_cs", 2), ("lsd_e", 3)):     # per-cluster candidate sweep (lsd_kernel<.., HI>)
            dec = qb.SlidingWindowDecoder(c, hz.shape[0], 5, 3, precision=prec, **dict(base, osd_method=m_, osd_order=o_, max_iter=3))
            print(m_, o_, prec, "logical errors", int(np.any((obs - dec.decode(det)) % 2, axis=1).sum()), flush=True)
    mc = qb.MonteCarlo(c, hz.shape[0], 5, 3, capacity=192, **base)
    print("fused", mc.run(n(500), 9)[0][0], flush=True)
    c3 = circuit("bb144_r10_p3e-3")                      # OSD-heavy: second tier and overflow route
    d3, o3 = qb.get_stim_mem_result(c3, n(400), seed=4)
    print("p=3e-3", int(np.any((o3 - qb.sliding_window_bposd_circuit_mem(d3, c3, hz, lz, 5, 3, **base)) % 2, axis=1).sum()), flush=True)
    hg = circuit("hgp225_r3_p1e-2")                      # whole-history window, generic paths
    dh, oh = qb.get_stim_mem_result(hg, n(100), seed=5)
    hzh, lzh = np.zeros((108, 225), dtype=np.uint8), np.zeros((9, 225), dtype=np.uint8)
    print("hgp", qb.sliding_window_bposd_circuit_mem(dh, hg, hzh, lzh, 3, 2, **base).sum(), flush=True)
    from scipy.sparse import csc_matrix
    rng = np.random.RandomState(2)                       # dense LSD merging on a tiny matrix (operation array compaction), and a
    for rows, cols, rate, cls, kwx in ((30, 90, 0.3, qb.BpLsdDecoder, dict(lsd_order=0)),            # tall matrix: LSD / OSD slab kernels
                                       (30, 90, 0.3, qb.BpLsdDecoder, dict(lsd_method="lsd_cs", lsd_order=2)),
                                       (1100, 2600, 0.02, qb.BpLsdDecoder, dict(lsd_order=0)),
                                       (1100, 2600, 0.02, qb.BpLsdDecoder, dict(lsd_method="lsd_e", lsd_order=2)),
                                       (1100, 2600, 0.02, qb.BpOsdDecoder, dict(osd_method="osd_0")),
                                       (1100, 2600, 0.004, qb.BpOsdDecoder, dict(osd_method="off", schedule="serial"))):
        indptr, indices = [0], []
        for j in range(cols):
            indices += list(np.sort(rng.choice(rows, size=3, replace=False)))
            indptr.append(len(indices))
        H = csc_matrix((np.ones(len(indices), dtype=np.uint8), np.array(indices), np.array(indptr)), shape=(rows, cols))
        err = (rng.rand(n(48), cols) < rate).astype(np.uint8)
        syn = np.asarray((H @ err.T).T % 2, dtype=np.uint8)
        kw2 = dict(max_iter=2, bp_method="minimum_sum", schedule="parallel")
        kw2.update(kwx)
        e = cls(H, error_rate=0.02, **kw2).decode_batch(syn)[0]
        print(cls.__name__, rows, cols, kwx, "weight", int(e.sum()), flush=True)
    # round-2 kernels: serial slab kernel with 8 / 16 lanes per column, product-sum in the generic flooding kernel, higher-order OSD
    # with the row transformation in a global slab, the device fan-out (three contexts on one GPU)
    for rows, cols, cw, kwx in ((300, 900, 9, dict(osd_method="off", schedule="serial")),
                                (300, 900, 9, dict(osd_method="off", schedule="serial", bp_method="product_sum")),
                                (300, 900, 15, dict(osd_method="off", schedule="serial")),
                                (300, 900, 9, dict(osd_method="off", bp_method="product_sum")),
                                (1100, 2600, 3, dict(osd_method="osd_cs", osd_order=1))):
        indptr, indices = [0], []
        for j in range(cols):
            indices += list(np.sort(rng.choice(rows, size=cw, replace=False)))
            indptr.append(len(indices))
        H = csc_matrix((np.ones(len(indices), dtype=np.uint8), np.array(indices), np.array(indptr)), shape=(rows, cols))
        err = (rng.rand(n(24), cols) < 0.01).astype(np.uint8)
        syn = np.asarray((H @ err.T).T % 2, dtype=np.uint8)
        kw2 = dict(max_iter=3, bp_method="minimum_sum", schedule="parallel")
        kw2.update(kwx)
        e = qb.BpOsdDecoder(H, error_rate=0.01, **kw2).decode_batch(syn)[0]
        print("round2", rows, cols, cw, kwx, "weight", int(e.sum()), flush=True)
    qb.set_devices([0, 0, 0])
    qb.devices.MIN_SHOTS_PER_DEVICE = 64
    det3, obs3 = qb.get_stim_mem_result(c, n(300), seed=3)
    print("fan-out", np.array_equal(det3, det), int(qb.sliding_window_bposd_circuit_mem(det3, c, hz, lz, 5, 3, **base).sum()), flush=True)
    qb.set_devices(None)
    rng = np.random.RandomState(1)
    print("phenom", qb.sliding_window_bposd_phenom_mem(rng.rand(n(64), 72 * 12) < 0.05, (rng.rand(72, 144) < 0.04).astype(int),
                                                       (rng.rand(12, 144) < 0.3).astype(int), 5, 3, error_rate=0.02, **base).sum(), flush=True)


if __name__ == "__main__":
    main()
