import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _ensure_built():
    """Compile libquits_b200.so (nvcc, sm_100a) when it is absent or stale; building is not a fallback, importing
    quits_b200 without the library still fails loudly."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_qb_build", os.path.join(ROOT, "quits_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.build()
    except RuntimeError:
        if not os.path.exists(mod.SO):
            raise


_ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def circuit_text(name):
    path = os.path.join(GOLDEN, "circuits", name + ".stim")
    if not os.path.exists(path) and os.path.exists(path + ".gz"):       # the large ones are committed compressed
        import gzip
        with gzip.open(path + ".gz", "rb") as f:
            return f.read().decode()
    with open(path) as f:
        return f.read()


def circuit_meta(name):
    with open(os.path.join(GOLDEN, "circuits", name + ".json")) as f:
        meta = json.load(f)
    hz = np.zeros(meta["hz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["hz_rows"]):
        hz[i, r] = 1
    lz = np.zeros(meta["lz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["lz_rows"]):
        lz[i, r] = 1
    return meta, hz, lz


def decode_case(case):
    """Fixture written by tools/make_golden.py: dict with det/obs bool arrays, pred_f32/pred_f64, seed, W, F, m."""
    z = np.load(os.path.join(GOLDEN, "decode", case + ".npz"))
    D, K, shots = int(z["D"]), int(z["K"]), int(z["shots"])
    det = np.unpackbits(z["det"], axis=1, bitorder="little")[:, :D].astype(np.bool_)
    obs = np.unpackbits(z["obs"], axis=1, bitorder="little")[:, :K].astype(np.bool_)
    return {"det": det, "obs": obs, "pred_f32": z["pred_f32"].astype(np.int64), "pred_f64": z["pred_f64"].astype(np.int64),
            "seed": int(z["seed"]), "shots": shots, "W": int(z["W"]), "F": int(z["F"]), "m": int(z["m"]), "D": D, "K": K}


DECODE_CASES = sorted(f[:-4] for f in os.listdir(os.path.join(GOLDEN, "decode")) if f.endswith(".npz")) \
    if os.path.isdir(os.path.join(GOLDEN, "decode")) else []


def case_circuit(case):
    return case.rsplit("_W", 1)[0]


BP_KW = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import cref
    cref.build()
    return cref


def random_circuit_text(rng, zbasis=None):
    """A random circuit in the grammar the reference's emitters produce (circuit.py:58-279) -- resets, H / CX layers with noise,
    measurements with detectors on earlier records, nested REPEAT blocks, comments, blank lines, empty-target noise lines."""
    nq = int(rng.integers(3, 10))
    lines, n_meas, n_obs = [], [0], 0
    draw = rng.random() < 0.7             # only Z-basis resets / measurements and CX: every detector is deterministic
    zbasis = draw if zbasis is None else zbasis

    def qubits(k=None):
        k = int(rng.integers(1, nq + 1)) if k is None else k
        return [int(x) for x in rng.choice(nq, size=min(k, nq), replace=False)]

    def prob():
        return float(rng.choice([0.001, 0.0123456789, 0.05, 0.2]))

    def emit(indent, depth):
        pad = "    " * indent
        for _ in range(int(rng.integers(3, 9))):
            kind = int(rng.integers(0, 10))
            if kind == 0:
                lines.append(pad + "%s %s" % ("R" if zbasis else rng.choice(["R", "RX"]), " ".join(map(str, qubits()))))
                lines.append(pad + "%s(%.10f) %s" % (rng.choice(["X_ERROR", "Z_ERROR"]), prob(), " ".join(map(str, qubits()))))
            elif kind == 1:
                qs = qubits()
                lines.append(pad + "H " + " ".join(map(str, qs)))
                if zbasis:
                    lines.append(pad + "H " + " ".join(map(str, qs)))        # back to the Z basis
                lines.append(pad + "DEPOLARIZE1(%.10f) %s" % (prob(), " ".join(map(str, qubits()))))
            elif kind in (2, 3):
                qs = qubits(2 * int(rng.integers(1, nq // 2 + 1)))
                qs = qs[:len(qs) // 2 * 2]
                if qs:
                    lines.append(pad + "CX " + " ".join(map(str, qs)))
                    lines.append(pad + "DEPOLARIZE2(%.10f) %s" % (prob(), " ".join(map(str, qs))))
            elif kind in (4, 5):
                qs = qubits()
                name = str(rng.choice(["M", "MR"])) if zbasis else str(rng.choice(["M", "MX", "MR"]))
                lines.append(pad + "X_ERROR(%.10f) %s" % (prob(), " ".join(map(str, qs))))
                lines.append(pad + name + " " + " ".join(map(str, qs)))
                n_meas[0] += len(qs)
                for _ in range(int(rng.integers(1, 3))):
                    look = sorted(set(int(x) for x in rng.integers(1, min(n_meas[0], 6) + 1, size=int(rng.integers(1, 4)))))
                    lines.append(pad + "DETECTOR " + " ".join("rec[-%d]" % k for k in look))
            elif kind == 6:
                lines.append(pad + "TICK")
                lines.append("")
                lines.append(pad + "# a comment")
            elif kind == 7:
                lines.append(pad + "DEPOLARIZE1(%.10f)" % prob())           # empty target list (circuit.py:106-124 on an empty layer)
            elif kind == 8 and depth < 2 and n_meas[0] >= 1:
                lines.append(pad + "REPEAT %d {" % int(rng.integers(1, 4)))
                emit(indent + 1, depth + 1)
                lines.append(pad + "}")
            elif kind == 9 and n_meas[0] >= 1:
                nonlocal_obs = int(rng.integers(0, 3))
                lines.append(pad + "OBSERVABLE_INCLUDE(%d) %s" % (nonlocal_obs, " ".join("rec[-%d]" % int(k) for k in rng.integers(1, min(n_meas[0], 4) + 1, size=2))))

    lines.append("R " + " ".join(map(str, range(nq))))
    emit(0, 0)
    lines.append("M " + " ".join(map(str, range(nq))))
    n_meas[0] += nq
    lines.append("DETECTOR rec[-1] rec[-%d]" % nq)
    lines.append("OBSERVABLE_INCLUDE(0) rec[-1]")
    return "\n".join(lines) + "\n"
