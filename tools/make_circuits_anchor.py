#!/usr/bin/env python
"""Freeze the circuits of the reference's end-to-end notebooks (doc/06A_end_to_end_demo_hgp.ipynb, doc/06B_end_to_end_demo_bb.ipynb)
at the noise rates where the notebooks print a non-zero logical failure rate: the only external numbers that exist for this path.
Built by the UNMODIFIED reference builders, called exactly as the notebooks call them.  Build container only.

    PYTHONHASHSEED=0 python tools/make_circuits_anchor.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from oracle import shims  # noqa: E402

shims.install()
sys.path.insert(0, "/root/reference/src")
from quits import ErrorModel, CircuitBuildOptions  # noqa: E402
from quits.qldpc_code import BbCode, HgpCode  # noqa: E402
from make_circuits import dump  # noqa: E402


def main():
    h = np.loadtxt("/root/reference/parity_check_matrices/n=12_dv=3_dc=4_dist=6.txt", dtype=int)
    hgp = HgpCode(h, h)
    for p, tag in ((2e-3, "hgp225_r15_p2e-3"),):
        circ = hgp.build_circuit(error_model=ErrorModel(p, p, p, p), num_rounds=15, basis="Z", circuit_build_options=CircuitBuildOptions(), seed=1)
        dump(tag, hgp, circ, {"code": "HgpCode(h,h) n=12_dv=3_dc=4_dist=6", "notebook": "doc/06A cell 4", "seed": 1, "rounds": 15, "p": p,
                              "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    bb = BbCode(l=15, m=3, A_x_pows=[9], A_y_pows=[1, 2], B_x_pows=[2, 7], B_y_pows=[0])
    for p, tag in ((1e-3, "bb90_r15_p1e-3"), (2e-3, "bb90_r15_p2e-3")):
        circ = bb.build_circuit(error_model=ErrorModel(p, p, p, p), num_rounds=15, basis="Z", circuit_build_options=CircuitBuildOptions())
        dump(tag, bb, circ, {"code": "BbCode(15,3,[9],[1,2],[2,7],[0])", "notebook": "doc/06B cell 4", "rounds": 15, "p": p, "basis": "Z"})


if __name__ == "__main__":
    main()
