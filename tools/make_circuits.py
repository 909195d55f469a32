#!/usr/bin/env python
"""Freeze the circuit texts the tests / bench use, by running the UNMODIFIED reference builders.

Run in the build container only (needs /root/reference; PYTHONHASHSEED=0 for the colouring-based
strategies).  Output: tests/golden/circuits/*.stim (+ *.json with hz/lz and
the build parameters).  The GPU box never runs this; it reads the committed fixtures.

    PYTHONHASHSEED=0 python tools/make_circuits.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

shims.install()
sys.path.insert(0, "/root/reference/src")
from quits import ErrorModel, CircuitBuildOptions  # noqa: E402
from quits.qldpc_code import BbCode, HgpCode  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "circuits")
os.makedirs(OUT, exist_ok=True)


def dump(name, code, circuit, meta):
    with open(os.path.join(OUT, name + ".stim"), "w") as f:
        f.write(circuit.text)
    meta = dict(meta)
    meta["hz_rows"] = [np.flatnonzero(r).tolist() for r in code.hz]
    meta["hz_shape"] = list(code.hz.shape)
    meta["lz_rows"] = [np.flatnonzero(r).tolist() for r in code.lz]
    meta["lz_shape"] = list(code.lz.shape)
    meta["n_lines"] = circuit.text.count("\n")
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, "qubits", circuit.num_qubits, "D", circuit.num_detectors, "K", circuit.num_observables,
          "meas", circuit.num_measurements)


def bb(l, m):
    return BbCode(l=l, m=m, A_x_pows=[3], A_y_pows=[1, 2], B_x_pows=[1, 2], B_y_pows=[3])


def main():
    opts = CircuitBuildOptions()
    # cfg 2: [[72,12,6]] BB, custom, 6 rounds, p=1e-3   (+ the 15-round text of doc/02A as parser KAT)
    c72 = bb(6, 6)
    for rounds, p, tag in ((6, 1e-3, "bb72_r6_p1e-3"), (15, 1e-3, "bb72_r15_p1e-3"), (6, 3e-3, "bb72_r6_p3e-3")):
        circ = c72.build_circuit(strategy="custom", error_model=ErrorModel(p, p, p, p), num_rounds=rounds, basis="Z",
                                 circuit_build_options=opts)
        dump(tag, c72, circ, {"code": "BbCode(6,6,[3],[1,2],[1,2],[3])", "strategy": "custom", "rounds": rounds, "p": p,
                               "basis": "Z"})
    # cfg 3: [[144,12,12]] gross code, custom, 10 rounds, p in {3e-3,1e-3,3e-4}
    c144 = bb(12, 6)
    for p, tag in ((1e-3, "bb144_r10_p1e-3"), (3e-3, "bb144_r10_p3e-3"), (3e-4, "bb144_r10_p3e-4")):
        circ = c144.build_circuit(strategy="custom", error_model=ErrorModel(p, p, p, p), num_rounds=10, basis="Z",
                                  circuit_build_options=opts)
        dump(tag, c144, circ, {"code": "BbCode(12,6,[3],[1,2],[1,2],[3])", "strategy": "custom", "rounds": 10, "p": p,
                                "basis": "Z"})
    # X-basis variant (exercises RX / MX)
    circ = c72.build_circuit(strategy="custom", error_model=ErrorModel(1e-3, 1e-3, 1e-3, 1e-3), num_rounds=3, basis="X",
                             circuit_build_options=opts)
    c72x = type("X", (), {"hz": c72.hx, "lz": c72.lx})
    dump("bb72_r3_p1e-3_X", c72x, circ, {"code": "BbCode(6,6,...)", "strategy": "custom", "rounds": 3, "p": 1e-3, "basis": "X"})
    # cfg 1: HGP from n=12 dv=3 dc=4, cardinal seed=1, 3 rounds, p=1e-2  (text is hash-seed dependent -> frozen)
    h = np.loadtxt("/root/reference/parity_check_matrices/n=12_dv=3_dc=4_dist=6.txt", dtype=int)
    hgp = HgpCode(h, h)
    circ = hgp.build_circuit(strategy="cardinal", error_model=ErrorModel(1e-2, 1e-2, 1e-2, 1e-2), num_rounds=3, basis="Z",
                             circuit_build_options=opts, seed=1)
    dump("hgp225_r3_p1e-2", hgp, circ, {"code": "HgpCode(h,h) n=12_dv=3_dc=4_dist=6", "strategy": "cardinal", "seed": 1,
                                         "rounds": 3, "p": 1e-2, "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    # zxcoloration skeleton (separate R / M per check type instead of MR): small toric HGP
    rep = np.array([[1, 1, 0], [0, 1, 1], [1, 0, 1]])
    tor = HgpCode(rep, rep)
    circ = tor.build_circuit(strategy="zxcoloration", error_model=ErrorModel(1e-3, 1e-3, 1e-3, 1e-3), num_rounds=3,
                             basis="Z", circuit_build_options=opts)
    dump("toric3_zxcol_r3_p1e-3", tor, circ, {"code": "HgpCode(rep3,rep3)", "strategy": "zxcoloration", "rounds": 3,
                                               "p": 1e-3, "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})


if __name__ == "__main__":
    main()
