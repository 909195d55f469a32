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
