#!/usr/bin/env python
"""BASELINE config 5: the [[1020,136]] quasi-cyclic lifted-product code (doc/01A_codes_basics.ipynb, QLP table, lift size 30),
zxcoloration circuit, 20 rounds, p = 5e-4 -- frozen from the UNMODIFIED reference builders (build container only):

    PYTHONHASHSEED=0 python tools/make_circuit_qlp1020.py

The Stim text is a few MB and is committed gzip-compressed (tests/golden/circuits/qlp1020_zxcol_r20_p5e-4.stim.gz).
"""
import gzip
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.make_circuits import OUT  # noqa: E402  (installs the shims, puts the reference on sys.path)
from quits import ErrorModel, CircuitBuildOptions  # noqa: E402
from quits.qldpc_code import QlpCode  # noqa: E402


def main():
    t0 = time.time()
    b = np.array([[0, 0, 0, 0, 0], [0, 2, 14, 24, 25], [0, 16, 11, 14, 13]])
    code = QlpCode(b, b, 30)
    p = 5e-4
    circ = code.build_circuit(strategy="zxcoloration", error_model=ErrorModel(p, p, p, p), num_rounds=20, basis="Z",
                              circuit_build_options=CircuitBuildOptions())
    name = "qlp1020_zxcol_r20_p5e-4"
    text = circ.text
    with gzip.open(os.path.join(OUT, name + ".stim.gz"), "wb", compresslevel=9) as f:
        f.write(text.encode())
    meta = {"code": "QlpCode(b, b, 30), b = doc/01A QLP table row lift 30", "strategy": "zxcoloration", "rounds": 20, "p": p, "basis": "Z",
            "PYTHONHASHSEED": os.environ.get("PYTHONHASHSEED"), "hz_shape": list(code.hz.shape), "lz_shape": list(code.lz.shape),
            "hz_rows": [np.flatnonzero(r).tolist() for r in code.hz], "lz_rows": [np.flatnonzero(r).tolist() for r in code.lz],
            "n_lines": text.count("\n"), "text_bytes": len(text)}
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(meta, f)
    print(name, "qubits", circ.num_qubits, "D", circ.num_detectors, "K", circ.num_observables, "text bytes", len(text),
          "in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    main()
