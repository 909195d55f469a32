#!/usr/bin/env python
"""More frozen circuit texts (same rules as tools/make_circuits.py: UNMODIFIED reference builders, build container only,
PYTHONHASHSEED=0 because the colouring-based strategies depend on set iteration order -- SURVEY.md section 0.6).

    PYTHONHASHSEED=0 python tools/make_circuits_more.py

  qt633_zxcol_r12_p1e-3   BASELINE config 4: the shipped 633 quantum-Tanner Hx/Hz pair through QldpcCode.from_parity_checks
                          (doc/01B_make_my_own_code.ipynb cells 3-5), zxcoloration circuit, 12 rounds
  bpc90_card_r10_p5e-4    the reference's own sliding-window tests (tests/test_sliding_window.py): BPC code, cardinal circuit seed 1
  hgp225_r15_p1e-3        the notebooks' HGP run (doc/03, doc/06A): cardinal circuit seed 1, 15 rounds -- windows of
                          540 x 6480, the largest the shared-memory kernels take
"""
import os
import sys

import numpy as np
from scipy.io import mmread

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.make_circuits import dump  # noqa: E402  (installs the shims, puts the reference on sys.path)
from quits import ErrorModel, CircuitBuildOptions  # noqa: E402
from quits.qldpc_code import HgpCode, QldpcCode  # noqa: E402

PCM = "/root/reference/parity_check_matrices/"


def main():
    opts = CircuitBuildOptions()
    stem = PCM + "633__C2xC2_AAp0_0_0_0_1_2_3_BBp0_0_0_1_1_2_2_k12_d11__"
    hz = (np.asarray(mmread(stem + "Hz.mtx").todense()).astype(int) & 1)
    hx = (np.asarray(mmread(stem + "Hx.mtx").todense()).astype(int) & 1)
    code = QldpcCode.from_parity_checks(hz, hx, compute_logicals=True)
    p = 1e-3
    circ = code.build_circuit(strategy="zxcoloration", error_model=ErrorModel(p, p, p, p), num_rounds=12, basis="Z",
                              circuit_build_options=opts)
    dump("qt633_zxcol_r12_p1e-3", code, circ, {"code": "QldpcCode.from_parity_checks(633 QTanner Hz, Hx)", "strategy": "zxcoloration",
                                               "rounds": 12, "p": p, "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    h = np.loadtxt(PCM + "n=12_dv=3_dc=4_dist=6.txt", dtype=int)
    hgp = HgpCode(h, h)
    circ = hgp.build_circuit(strategy="cardinal", error_model=ErrorModel(p, p, p, p), num_rounds=15, basis="Z",
                             circuit_build_options=opts, seed=1)
    dump("hgp225_r15_p1e-3", hgp, circ, {"code": "HgpCode(h,h) n=12_dv=3_dc=4_dist=6", "strategy": "cardinal", "seed": 1, "rounds": 15,
                                         "p": p, "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    # the code and circuit of the reference's own sliding-window tests (tests/test_sliding_window.py:10-33,47-52,106-111):
    # BpcCode([0,1,5], [0,8,13], 15, 3), cardinal circuit seed 1, p = 5e-4, 10 rounds
    from quits.qldpc_code import BpcCode
    bpc = BpcCode([0, 1, 5], [0, 8, 13], 15, 3)
    bpc.build_circuit(strategy="cardinal", seed=1)
    p = 5e-4
    circ = bpc.build_circuit(strategy="cardinal", error_model=ErrorModel(p, p, p, p), num_rounds=10, basis="Z", seed=1)
    dump("bpc90_card_r10_p5e-4", bpc, circ, {"code": "BpcCode([0,1,5],[0,8,13],15,3)", "strategy": "cardinal", "seed": 1, "rounds": 10, "p": p,
                                             "basis": "Z", "depth": int(bpc.depth), "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    # the other code families / strategies of the reference's tests (tests/test_codes.py:232-343), 6 rounds
    from quits.qldpc_code import LcsCode, QlpCode
    p = 1e-3
    lcs = LcsCode(5, 3)
    circ = lcs.build_circuit(strategy="cardinal", error_model=ErrorModel(p, p, p, p), num_rounds=6, basis="Z", seed=1)
    dump("lcs_card_r6_p1e-3", lcs, circ, {"code": "LcsCode(5, 3)", "strategy": "cardinal", "seed": 1, "rounds": 6, "p": p, "basis": "Z",
                                          "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    circ = bpc.build_circuit(strategy="cardinalNSmerge", error_model=ErrorModel(p, p, p, p), num_rounds=6, basis="Z", seed=1)
    dump("bpc90_nsmerge_r6_p1e-3", bpc, circ, {"code": "BpcCode([0,1,5],[0,8,13],15,3)", "strategy": "cardinalNSmerge", "seed": 1, "rounds": 6,
                                               "p": p, "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})
    p = 5e-4
    b = np.array([[0, 0, 0, 0, 0], [0, 2, 4, 7, 11], [0, 3, 10, 14, 15]])
    qlp = QlpCode(b, b, 16)
    circ = qlp.build_circuit(strategy="cardinal", error_model=ErrorModel(p, p, p, p), num_rounds=6, basis="Z", seed=1)
    # (committed gzip-compressed: gzip -9 tests/golden/circuits/qlp544_card_r6_p5e-4.stim)
    dump("qlp544_card_r6_p5e-4", qlp, circ, {"code": "QlpCode(b, b, 16), [[544,80]]", "strategy": "cardinal", "seed": 1, "rounds": 6, "p": p,
                                             "basis": "Z", "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED")})


if __name__ == "__main__":
    main()
