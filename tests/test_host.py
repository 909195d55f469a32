"""CPU tests of the product's host side (C++ behind the C ABI): Stim-text front end, DEM, check matrix, window plan,
error behaviour, and that the library exports what include/quits_b200.h declares.  No compute call needs a GPU here."""
import ctypes
import hashlib
import json
import os
import re
import warnings

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, case_circuit, circuit_meta, circuit_text, decode_case, random_circuit_text
from oracle import cref, dem as odem, shims, stimtext

import quits_b200 as qb
from quits_b200 import _native as N
from quits_b200.decoder.base import WindowPlan

CIRCUITS = sorted(f[:-5] if f.endswith(".stim") else f[:-8] for f in os.listdir(os.path.join(GOLDEN, "circuits"))
                  if f.endswith(".stim") or f.endswith(".stim.gz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_library_exports_every_declared_symbol():
    with open(os.path.join(ROOT, "include", "quits_b200.h")) as f:
        header = f.read()
    declared = sorted(set(re.findall(r"\b(qb_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(N.SO_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(N.SYMBOLS) == declared
    assert lib.qb_version() == 100


def test_no_gpu_means_loud_failure():
    if N.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(N.QbCudaError, match="no CPU fallback"):
        qb.get_stim_mem_result(circuit_text("toric3_zxcol_r3_p1e-3"), 10, seed=1)
    _, hz, lz = circuit_meta("toric3_zxcol_r3_p1e-3")
    with pytest.raises(N.QbCudaError):
        qb.sliding_window_bposd_circuit_mem(np.zeros((4, 45), dtype=bool), circuit_text("toric3_zxcol_r3_p1e-3"), hz, lz, 3, 2,
                                            bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")


@pytest.mark.parametrize("name", CIRCUITS)
def test_front_end_and_dem_match_oracle(name):
    text = circuit_text(name)
    c = qb.Circuit(text)
    fc = stimtext.parse_flat(text)
    assert (c.num_qubits, c.num_measurements, c.num_detectors, c.num_observables, c.num_flat_ops) == \
        (fc.n_qubits, fc.n_meas, fc.n_det, fc.n_obs, len(fc.ops))
    kind, arg, ts, tg = c.flat()
    ok, oa, ots, otg = cref.flat_arrays(fc)
    assert np.array_equal(kind, ok) and np.array_equal(arg, oa) and np.array_equal(ts, ots) and np.array_equal(tg, otg[:len(tg)])
    d = c.detector_error_model()
    od = odem.analyze(fc)
    e = d.errors()
    assert d.num_errors == len(od.probs)
    assert np.array_equal(e["probs"], np.array(od.probs))                   # bit-identical priors
    assert e["det_idx"].tolist() == [x for ds in od.dets for x in ds]
    assert e["obs_idx"].tolist() == [x for ds in od.obs for x in ds]
    assert np.array_equal(np.diff(e["det_ptr"]), [len(x) for x in od.dets])
    assert list(zip(e["rep_op"].tolist(), e["rep_tgt"].tolist(), e["rep_code"].tolist())) == [tuple(r) for r in od.rep]
    gpath = os.path.join(GOLDEN, "dem", name + ".json")
    if os.path.exists(gpath):
        with open(gpath) as f:
            g = json.load(f)
        assert d.num_errors == g["n_errors"] and sha(e["probs"]) == g["probs_sha256"] and sha(e["det_idx"]) == g["det_idx_sha256"]
        assert sha(e["obs_idx"]) == g["obs_idx_sha256"]


def test_canonical_stim_printing_is_accepted():
    """What str(stim.Circuit) would hand us: fused target lists, shortest-repr numbers, indented REPEAT bodies."""
    text = circuit_text("bb72_r6_p1e-3")
    a = qb.Circuit(text)
    b = qb.Circuit(stimtext.canonical_text(text))
    assert (a.num_qubits, a.num_measurements, a.num_detectors, a.num_observables) == \
        (b.num_qubits, b.num_measurements, b.num_detectors, b.num_observables)
    ea, eb = a.detector_error_model().errors(), b.detector_error_model().errors()
    assert np.array_equal(ea["probs"], eb["probs"]) and np.array_equal(ea["det_idx"], eb["det_idx"])


@pytest.mark.parametrize("case", sorted(f[:-5] for f in os.listdir(os.path.join(GOLDEN, "windows"))))
def test_window_plan_equals_reference_spacetime(case):
    """The digests were produced by the reference's own spacetime() (decoder/base.py:134-190), tools/make_golden.py."""
    with open(os.path.join(GOLDEN, "windows", case + ".json")) as f:
        g = json.load(f)
    c = qb.Circuit(circuit_text(g["circuit"]))
    plan = WindowPlan(c.detector_error_model(), g["m"], g["W"], g["F"])
    assert plan.n_windows == len(g["windows"])
    for k, e in enumerate(g["windows"]):
        w = plan.window(k)
        H, L = w["H"], w["L"]
        assert list(H.shape) == e["H_shape"] and H.nnz == e["H_nnz"]
        assert sha(H.indptr.astype(np.int64)) == e["H_indptr"] and sha(H.indices.astype(np.int32)) == e["H_indices"]
        assert list(L.shape) == e["L_shape"] and sha(L.indptr.astype(np.int64)) == e["L_indptr"]
        assert sha(L.indices.astype(np.int32)) == e["L_indices"]
        assert sha(w["priors"]) == e["priors"]
        if "U_shape" in e:
            U = w["U"]
            assert list(U.shape) == e["U_shape"] and sha(U.indptr.astype(np.int64)) == e["U_indptr"]
            assert sha(U.indices.astype(np.int32)) == e["U_indices"]
    # the spacetime() drop-in returns the same four lists
    _, hz, _ = circuit_meta(g["circuit"])
    checks, obs, priors, updates = qb.spacetime(c, hz, g["W"], g["F"], g["num_cor_rounds"])
    assert len(checks) == len(g["windows"]) and len(updates) == len(g["windows"]) - 1
    assert [list(h.shape) for h in checks] == [e["H_shape"] for e in g["windows"]]


def test_matrix_conversion_on_a_foreign_dem():
    """detector_error_model_to_matrix on a stim-shaped object that is not ours (the oracle's shim)."""
    text = circuit_text("toric3_zxcol_r3_p1e-3")
    theirs = shims.Circuit(text).detector_error_model()
    H1, L1, p1 = qb.detector_error_model_to_matrix(theirs)
    H2, L2, p2 = qb.detector_error_model_to_matrix(qb.Circuit(text).detector_error_model())
    assert (H1 != H2).nnz == 0 and (L1 != L2).nnz == 0 and np.array_equal(p1, p2)
    assert H1.dtype == np.uint8 and H1.shape == (45, 234)


def test_merge_rule_keys_on_detectors_only():
    """base.py:93-99: equal detector sets merge (XOR-combined probability), the first sighting's observables win."""
    class T:
        def __init__(self, v, o): self.val, self._o = v, o
        def is_relative_detector_id(self): return not self._o
        def is_logical_observable_id(self): return self._o
    class I:
        type = "error"
        def __init__(self, p, d, o): self._p, self._t = p, [T(x, False) for x in d] + [T(x, True) for x in o]
        def args_copy(self): return [self._p]
        def targets_copy(self): return self._t
    class Dem:
        num_detectors, num_observables = 3, 2
        def flattened(self): return [I(0.1, [0, 1], [0]), I(0.2, [2], []), I(0.3, [1, 0], [1])]
    H, L, p = qb.detector_error_model_to_matrix(Dem())
    assert H.shape == (3, 2) and L.shape == (2, 2)
    assert H.toarray().tolist() == [[1, 0], [1, 0], [0, 1]]
    assert L.toarray().tolist() == [[1, 0], [0, 0]]
    assert p[0] == 0.1 * (1 - 0.3) + 0.3 * (1 - 0.1) and p[1] == 0.2


def test_error_behaviour():
    text = circuit_text("toric3_zxcol_r3_p1e-3")
    _, hz, lz = circuit_meta("toric3_zxcol_r3_p1e-3")
    with pytest.raises(ValueError, match="F cannot be zero"):
        qb.spacetime(text, hz, 3, 0, 1)
    with pytest.raises(ZeroDivisionError):                 # the reference divides by F before spacetime() checks it
        qb.sliding_window_bposd_circuit_mem(np.zeros((2, 45), dtype=bool), text, hz, lz, 3, 0, bp_method="minimum_sum",
                                            schedule="parallel")
    with pytest.raises(NotImplementedError):
        qb.Circuit("PAULI_CHANNEL_1(0.1,0.1,0.1) 0\nM 0\n")
    for bad in ("FOO 1 2\n", "CX 0\n", "REPEAT 3 {\nH 0\n", "}\n", "M 0\nDETECTOR rec[-2]\n", "X_ERROR(0.7) 0\n", "H -1\n"):
        with pytest.raises(ValueError):
            qb.Circuit(bad)
    # a window without faults is a ValueError like in the reference (base.py:161-162)
    quiet = "R 0 1\nM 0 1\nDETECTOR rec[-1]\nDETECTOR rec[-2]\nM 0 1\nDETECTOR rec[-1] rec[-3]\nDETECTOR rec[-2] rec[-4]\n" \
            "M 0 1\nDETECTOR rec[-1] rec[-3]\nDETECTOR rec[-2] rec[-4]\n"
    with pytest.raises(ValueError):
        qb.spacetime(quiet, np.zeros((2, 2)), 2, 1, 1)
    # W larger than the history: the reference warns and decodes the whole history (sliding_window.py:140)
    if N.device_count() == 0:
        with warnings.catch_warnings(record=True) as rec:
            warnings.simplefilter("always")
            with pytest.raises(N.QbCudaError):
                qb.sliding_window_bposd_circuit_mem(np.zeros((2, 45), dtype=bool), text, hz, lz, 9, 2, bp_method="minimum_sum",
                                                    schedule="parallel")
        assert any("whole history" in str(w.message) for w in rec)


def test_repeat_and_split_semantics():
    """REPEAT unrolling, rec[-k] resolution, and sequential semantics inside one instruction (CX 0 1 1 2)."""
    c = qb.Circuit("R 0 1 2\nX_ERROR(0.25) 0\nCX 0 1 1 2\nREPEAT 2 {\n    M 2\n    DETECTOR rec[-1]\n}\nOBSERVABLE_INCLUDE(1) rec[-1] rec[-2]\n")
    assert (c.num_qubits, c.num_measurements, c.num_detectors, c.num_observables) == (3, 2, 2, 2)
    assert c.num_tape_ops >= c.num_flat_ops - 1        # the CX was split in two, the DETECTORs are separate blocks
    e = c.detector_error_model().errors()
    assert e["probs"].tolist() == [0.25] and e["det_idx"].tolist() == [0, 1] and e["obs_idx"].tolist() == []


def test_shard_ranges():
    for total in (0, 1, 63, 64, 1000, 10**7 + 3):
        for world in (1, 2, 3, 8):
            spans = [qb.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c2, d) in zip(spans, spans[1:]):
                assert b == c2 and a <= b
            assert all(lo % 64 == 0 for lo, _ in spans)


@pytest.mark.parametrize("precision", [64, 32])
def test_bp_layout_is_a_valid_conflict_free_arrangement(precision):
    """layout.cpp: the record order is a permutation sorted by column weight, every row's slots are a permutation of
    0..len-1, and the predicted shared-memory passes of the bit sweep are within 20 % of the conflict-free minimum
    (the naive layout needs 1.7-2.8x)."""
    circuit = qb.Circuit(circuit_text("bb144_r10_p1e-3"))
    plan = WindowPlan(circuit.detector_error_model(), 72, 5, 3)
    for k in (1, 3):
        H = plan.window(k)["H"]
        order = np.zeros(H.shape[1], dtype=np.int32)
        slot = np.zeros(H.nnz, dtype=np.int32)
        ratios = np.zeros(4)
        N.check(N.lib().qb_plan_layout(plan._h, k, precision, N.ptr(order), N.ptr(slot), N.ptr(ratios)))
        assert sorted(order.tolist()) == list(range(H.shape[1]))
        wts = np.diff(H.indptr)[order]
        assert np.all(wts[:-1] >= wts[1:])
        for i in range(H.shape[0]):
            s = slot[H.indices == i]
            assert sorted(s.tolist()) == list(range(len(s)))
        assert ratios[0] > 2.0 and ratios[1] > 1.5
        assert ratios[2] < 1.1 and ratios[3] < 1.25, ratios


def test_phenomenological_windows_equal_the_reference_matrices():
    """_phenom_plan builds exactly the matrices of reference sliding_window.py:56-69 and the commit / carry rule of :86-88."""
    from quits_b200.decoder.sliding_window import _phenom_plan
    rng = np.random.RandomState(7)
    hz = (rng.rand(5, 9) < 0.4).astype(int)
    hz[:, 0] = 1
    lz = (rng.rand(2, 9) < 0.5).astype(int)
    m, n = hz.shape
    W, F, rounds = 4, 2, 6
    ncor = (2 + rounds - W + F - 1) // F
    W_last = rounds + 2 - F * ncor
    plan = _phenom_plan(hz, lz, W, F, ncor, W_last, m * (rounds + 2), 0.03, 0.04)
    assert plan.n_windows == ncor + 1
    for k in range(plan.n_windows):
        last = k == ncor
        Wk = W_last if last else W
        B = np.eye(Wk, dtype=int)
        for i in range(1, Wk):
            B[i, i - 1] = 1
        if last:
            B = B[:, :Wk - 1]
        ref = np.column_stack((np.kron(np.eye(Wk, dtype=int), hz), np.kron(B, np.eye(m, dtype=int))))
        w = plan.window(k)
        assert w["row0"] == F * k * m and np.array_equal(w["H"].toarray(), ref)
        assert np.allclose(w["priors"], 0.04 if last else 0.03)
        e = rng.randint(0, 2, size=ref.shape[1])
        blocks = Wk if last else F
        corr = e[:blocks * n].reshape(blocks, n).sum(axis=0) % 2
        assert np.array_equal(w["L"].toarray() @ e[:w["ncommit"]] % 2, lz @ corr % 2)
        if not last:
            assert np.array_equal(w["U"].toarray() @ e % 2, e[W * n + (F - 1) * m:W * n + F * m])
    with pytest.raises(ValueError):
        qb.sliding_window_bposd_phenom_mem(np.zeros((1, m * 8), dtype=bool), hz, lz, 4, 0, 0.05)


def test_explicit_plan_validation():
    """qb_plan_create_explicit rejects malformed windows with the boundary's exception types."""
    from scipy.sparse import csc_matrix
    m, K, D = 2, 1, 6
    H = csc_matrix(np.array([[1, 0, 1], [0, 1, 1], [1, 1, 0], [0, 0, 1]], dtype=np.uint8))
    L = csc_matrix(np.array([[1, 0, 1]], dtype=np.uint8))
    U = csc_matrix(np.array([[0, 1, 0], [0, 0, 1]], dtype=np.uint8))
    good = [{"row0": 0, "H": H, "priors": [0.1, 0.2, 0.3], "L": L, "U": U},
            {"row0": 2, "H": H, "priors": [0.1, 0.2, 0.3], "L": L, "U": None}]
    plan = WindowPlan.explicit(m, K, D, good)
    assert plan.n_windows == 2 and plan.K == 1 and plan.D == 6
    w0 = plan.window(0)
    assert np.array_equal(w0["H"].toarray(), H.toarray()) and np.array_equal(w0["U"].toarray(), U.toarray())
    with pytest.raises(TypeError):                      # rows outside the detector range
        WindowPlan.explicit(m, K, D, [dict(good[0], row0=4), good[1]])
    with pytest.raises(TypeError):                      # the last window must not carry, the others must carry m rows
        WindowPlan.explicit(m, K, D, [good[0], dict(good[1], U=U)])
    with pytest.raises(TypeError):
        WindowPlan.explicit(m, K, D, [dict(good[0], U=None), good[1]])
    with pytest.raises(TypeError):                      # observable index out of range
        WindowPlan.explicit(m, K, D, [dict(good[0], L=csc_matrix(np.array([[1, 0, 0], [0, 1, 0]], dtype=np.uint8))), good[1]])


def test_phenom_argument_errors():
    hz = np.eye(3, 4, dtype=int)
    lz = np.ones((1, 4), dtype=int)
    det = np.zeros((2, 3 * 6), dtype=bool)
    with pytest.raises(ValueError):
        qb.sliding_window_bposd_phenom_mem(det, hz, lz, 3, 0, 0.05)
    with pytest.raises(ValueError):                     # F > W: the reference fails reshaping F blocks out of W
        qb.sliding_window_bposd_phenom_mem(det, hz, lz, 2, 3, 0.05, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")
    # the rate is mandatory, as in the reference (decoder/bposd.py:33-36, bplsd.py:33-36); error_rate is the deprecated alias
    for fn in (qb.sliding_window_bposd_phenom_mem, qb.sliding_window_bplsd_phenom_mem):
        with pytest.raises(ValueError, match="eff_error_rate_per_fault"):
            fn(det, hz, lz, 3, 2)
        with pytest.raises(ValueError, match="cannot be zero"):
            fn(det, hz, lz, 3, 0, error_rate=0.05)
        with pytest.raises(ValueError, match="cannot be zero"):
            fn(det, hz, lz, 3, 0, eff_error_rate_per_fault=0.05)
    # a foreign inner decoder class goes through the plug-in loop (one object per window, one call per shot and window)
    class Zero:
        def __init__(self, pcm, error_rate=None):
            self.n = pcm.shape[1]

        def decode(self, s):
            return np.zeros(self.n, dtype=int)
    pred = qb.sliding_window_phenom_mem(det, hz, lz, 3, 2, Zero, Zero, {"error_rate": 0.1}, {"error_rate": 0.1}, "decode", "decode")
    assert pred.shape == (2, 1) and pred.dtype == np.int64 and not pred.any()


def test_lsd_option_mapping():
    """BpLsdDecoder keywords (reference decoder/bplsd.py:38-49,74-83): order 0 maps onto the engine's lsd_0 whatever lsd_method says;
    beyond order 0 lsd_cs / lsd_e select the per-cluster candidate sweep with lsd_order as its order."""
    from quits_b200.decoder.inner import lsd_engine_options
    from quits_b200.engine import bp_options
    kw = lsd_engine_options({"bp_method": "product_sum", "max_iter": 2, "schedule": "serial", "lsd_method": "lsd_cs", "lsd_order": 0})
    assert kw == {"bp_method": "product_sum", "max_iter": 2, "schedule": "serial", "osd_method": "lsd_0", "osd_order": 0}
    o = bp_options(**kw)
    assert (o.osd_method, o.osd_order, o.bp_method, o.schedule) == (3, 0, 1, 1)
    assert lsd_engine_options({"lsd_method": "off"})["osd_method"] == "off"
    o = bp_options(**lsd_engine_options({"lsd_method": "lsd_cs", "lsd_order": 1}))          # the reference's own LSD test: order 1
    assert (o.osd_method, o.osd_order) == (5, 1)
    o = bp_options(**lsd_engine_options({"lsd_method": "lsd_e", "lsd_order": 3}))
    assert (o.osd_method, o.osd_order) == (4, 3)
    o = bp_options(**lsd_engine_options({"lsd_method": "lsd_0", "lsd_order": 7}))            # lsd_0 is order 0 whatever lsd_order says
    assert (o.osd_method, o.osd_order) == (3, 0)
    with pytest.raises(ValueError):
        lsd_engine_options({"lsd_method": "osd_cs"})
    with pytest.raises(ValueError):
        lsd_engine_options({"lsd_method": "lsd_cs", "lsd_order": -1})
    with pytest.raises(NotImplementedError):
        lsd_engine_options({"lsd_method": "lsd_cs", "lsd_order": 1, "bits_per_step": 2})


def test_drop_in_signatures_equal_the_reference():
    """Parameter names, order and defaults of every drop-in function equal the reference's (tests/golden/ref_signatures.json was
    written with inspect.signature over quits.decoder / quits.simulation of the reference tree, v1.1.0)."""
    import inspect
    import json
    with open(os.path.join(GOLDEN, "ref_signatures.json")) as f:
        want = json.load(f)
    assert len(want) == 10
    for name, params in want.items():
        got = [[k, repr(v.default) if v.default is not inspect._empty else None]
               for k, v in inspect.signature(getattr(qb, name)).parameters.items()]
        assert got == params, name


def test_baseline_config5_host_pipeline():
    """BASELINE config 5 -- QlpCode(b, b, 30) = [[1020,136]], zxcoloration, 20 rounds, p = 5e-4, frozen from the reference's own
    builders (tools/make_circuit_qlp1020.py): the C++ front end, DEM analyser and window planner reproduce the sizes SURVEY
    Appendix C probed with stim-shaped Python (9900 x 133320, nnz 659766, column weight <= 15, row weight <= 87, sum of priors
    159.86; seven W5/F3 windows), in well under a second where the Python stand-in needed 9 s + minutes of scipy slicing."""
    import time
    text = circuit_text("qlp1020_zxcol_r20_p5e-4")
    t0 = time.perf_counter()
    c = qb.Circuit(text)
    assert (c.num_qubits, c.num_detectors, c.num_observables) == (1920, 9900, 136)
    dem = c.detector_error_model()
    H, L, pri = qb.detector_error_model_to_matrix(dem)
    plan = WindowPlan(dem, 450, 5, 3)
    dt = time.perf_counter() - t0
    assert H.shape == (9900, 133320) and H.nnz == 659766 and L.shape == (136, 133320)
    assert int(np.diff(H.indptr).max()) == 15 and int(np.diff(H.tocsr().indptr).max()) == 87
    assert abs(float(pri.sum()) - 159.86) < 0.01
    assert plan.n_windows == 7
    shapes = [(plan.window(k)["H"].shape, plan.window(k)["H"].nnz, plan.window(k)["ncommit"]) for k in range(7)]
    assert shapes[0] == ((2250, 29250), 133984, 16650)
    assert all(sh == ((2250, 31500), 150830, 18900) for sh in shapes[1:6])
    assert shapes[6] == ((1800, 22170), 114184, 22170)
    assert plan.window(3)["U"].shape == (450, 18900)
    assert dt < 30.0


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/quits"), reason="needs the reference tree (build container only)")
def test_compat_runs_the_unmodified_reference_package():
    """quits_b200.compat.install() registers the engine under the names the reference imports (stim, ldpc.bposd_decoder,
    ldpc.bplsd_decoder); the UNMODIFIED reference package then builds its circuits into the engine's Circuit (C++ front end), walks
    it with its own helpers, and its own detector_error_model_to_matrix / spacetime run on the engine's DEM.  Run in a subprocess:
    the session's other tests keep the oracle's shims.  (No GPU: nothing is sampled or decoded here.)"""
    import subprocess
    import sys
    code = r'''
import sys, json, hashlib
import numpy as np
sys.path.insert(0, %r)
import quits_b200.compat as compat
assert compat.install(force=True)
sys.path.insert(0, "/root/reference/src")
import quits
from quits import ErrorModel, CircuitBuildOptions
from quits.qldpc_code import BbCode
from quits.circuit import check_overlapping_CX
from quits.decoder import spacetime, detector_error_model_to_matrix
import stim, ldpc.bposd_decoder, ldpc.bplsd_decoder
assert stim.Circuit is compat.StimCircuit and ldpc.bposd_decoder.BpOsdDecoder.__module__.startswith("quits_b200")
code = BbCode(l=6, m=6, A_x_pows=[3], A_y_pows=[1, 2], B_x_pows=[1, 2], B_y_pows=[3])
p = 1e-3
circ = code.build_circuit(strategy="custom", error_model=ErrorModel(p, p, p, p), num_rounds=6, basis="Z",
                          circuit_build_options=CircuitBuildOptions())
assert isinstance(circ, compat.StimCircuit)
assert check_overlapping_CX(circ, verbose=False) == []
H, L, pri = detector_error_model_to_matrix(circ.detector_error_model(decompose_errors=False))
checks, obs, priors, updates = spacetime(circ, code.hz, 5, 3, 1)
print(json.dumps({"text_sha": hashlib.sha256(circ.text.encode()).hexdigest(), "n_instr": len(circ), "H": list(H.shape), "nnz": int(H.nnz),
                  "sum": float(np.sum(pri)), "windows": [list(c.shape) for c in checks], "D": circ.num_detectors, "K": circ.num_observables}))
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads(out.stdout.strip().split("\n")[-1])
    want_text = circuit_text("bb72_r6_p1e-3")
    assert got["text_sha"] == hashlib.sha256(want_text.encode()).hexdigest()        # the frozen fixture IS what the reference builds
    assert got["H"] == [288, 2592] and got["nnz"] == 9036 and abs(got["sum"] - 4.69) < 0.01
    assert got["windows"] == [[180, 1764], [180, 1548]] and (got["D"], got["K"]) == (288, 12)
    assert got["n_instr"] > 10


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/quits"), reason="needs the reference tree (build container only)")
def test_every_code_family_of_the_reference_builds_into_the_engine():
    """The code families and circuit strategies of the reference's own tests (tests/test_codes.py:159-414: HGP cardinal and
    zxcoloration, QLP, BPC cardinal and cardinalNSmerge, LCS, BB custom), built by the UNMODIFIED reference on top of
    quits_b200.compat: every text is accepted by the C++ front end, gives D = m (rounds + 1) detectors... and K observables, a DEM
    whose matrix has the detectors' row count and no column heavier than the BP kernels take, and a W5/F3 window plan that covers
    every detector row exactly once in its committed rows."""
    import subprocess
    import sys
    code = r'''
import sys, json
import numpy as np
sys.path.insert(0, %r)
import quits_b200.compat as compat
compat.install(force=True)
sys.path.insert(0, "/root/reference/src")
from quits import ErrorModel
from quits.qldpc_code import BbCode, BpcCode, HgpCode, LcsCode, QlpCode
from quits.decoder import detector_error_model_to_matrix
from quits_b200.decoder.base import WindowPlan
h = np.loadtxt("/root/reference/parity_check_matrices/n=12_dv=3_dc=4_dist=6.txt", dtype=int)
b = np.array([[0, 0, 0, 0, 0], [0, 2, 4, 7, 11], [0, 3, 10, 14, 15]])
cases = [("hgp_cardinal", HgpCode(h, h), dict(strategy="cardinal", seed=1)), ("hgp_zxcol", HgpCode(h, h), dict(strategy="zxcoloration")),
         ("bpc_cardinal", BpcCode([0, 1, 5], [0, 8, 13], 15, 3), dict(strategy="cardinal", seed=1)),
         ("bpc_nsmerge", BpcCode([0, 1, 5], [0, 8, 13], 15, 3), dict(strategy="cardinalNSmerge", seed=1)),
         ("lcs_cardinal", LcsCode(5, 3), dict(strategy="cardinal", seed=1)),
         ("qlp_cardinal", QlpCode(b, b, 16), dict(strategy="cardinal", seed=1)),
         ("bb_custom", BbCode(l=6, m=6, A_x_pows=[3], A_y_pows=[1, 2], B_x_pows=[1, 2], B_y_pows=[3]), dict(strategy="custom"))]
out = {}
for name, code, kw in cases:
    p, rounds = 5e-4, 6
    circ = code.build_circuit(error_model=ErrorModel(p, p, p, p), num_rounds=rounds, basis="Z", **kw)
    m, K = code.hz.shape[0], code.lz.shape[0]
    H, L, pri = detector_error_model_to_matrix(circ.detector_error_model(decompose_errors=False))
    plan = WindowPlan(circ.detector_error_model(), m, 5, 3)
    rows = [plan.window(k)["rows"] for k in range(plan.n_windows)]
    out[name] = {"qubits": circ.num_qubits, "D": circ.num_detectors, "K": circ.num_observables, "m": m, "k": K, "rounds": rounds,
                 "H": list(H.shape), "colw": int(np.diff(H.indptr).max()), "priors_ok": bool(((pri > 0) & (pri < 0.5)).all()),
                 "windows": plan.n_windows, "rows": rows, "L_rows": int(L.shape[0])}
print(json.dumps(out))
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=dict(os.environ, PYTHONHASHSEED="0"))
    assert out.returncode == 0, out.stderr[-3000:]
    got = json.loads(out.stdout.strip().split("\n")[-1])
    assert len(got) == 7
    for name, g in got.items():
        assert g["D"] == g["m"] * (g["rounds"] + 2), name          # one detector layer per round, plus the first and the final data layer
        assert g["K"] == g["k"] == g["L_rows"], name
        assert g["H"][0] == g["D"] and g["colw"] <= 16 and g["priors_ok"], name
        assert g["windows"] == 2 and g["rows"] == [5 * g["m"], 5 * g["m"]], name       # rounds + 2 = 8 layers: W5/F3 -> layers 0-4, 3-7


def test_front_end_and_dem_match_oracle_on_random_circuits():
    """400 random circuits in the emitters' grammar (those whose random detectors are not deterministic must be refused by both
    sides with the same message): the C++ parser's flattened op list and the C++ backward DEM analyser equal the
    oracle's Python restatements item for item (probabilities bit-identical, stim's lexicographic order)."""
    rng = np.random.default_rng(2024)
    n_rep = n_err = n_nondet = 0
    for trial in range(400):
        text = random_circuit_text(rng)
        c = qb.Circuit(text)
        fc = stimtext.parse_flat(text)
        assert (c.num_qubits, c.num_measurements, c.num_detectors, c.num_observables, c.num_flat_ops) == \
            (fc.n_qubits, fc.n_meas, fc.n_det, fc.n_obs, len(fc.ops)), text
        kind, arg, ts, tg = c.flat()
        ok, oa, ots, otg = cref.flat_arrays(fc)
        assert np.array_equal(kind, ok) and np.array_equal(arg, oa) and np.array_equal(ts, ots) and np.array_equal(tg, otg[:len(tg)]), text
        n_rep += "REPEAT" in text
        try:
            od = odem.analyze(fc)
        except ValueError as err:                       # a detector that is not deterministic: both sides must say so, in the same words
            with pytest.raises(ValueError) as info:
                c.detector_error_model()
            assert str(info.value) == str(err), text
            n_nondet += 1
            continue
        e = c.detector_error_model().errors()
        assert np.array_equal(e["probs"], np.array(od.probs, dtype=np.float64)), text
        assert e["det_idx"].tolist() == [x for ds in od.dets for x in ds] and e["obs_idx"].tolist() == [x for ds in od.obs for x in ds], text
        n_err += len(od.probs)
    assert n_rep > 30 and n_err > 1000 and 10 < n_nondet < 300, (n_rep, n_err, n_nondet)


class _ForeignDecoder:
    """A decoder class this engine knows nothing about, with the call shape the reference's plug-in seam expects
    (decoder(pcm, **params) then a named decode method, sliding_window.py:146-153,171,182): the oracle's BP + OSD behind it."""

    def __init__(self, pcm, my_rates=None, **kw):
        from oracle import cref
        self._d = cref.BpOsd(pcm, my_rates, **kw)
        self.calls = 0

    def run_it(self, syndrome):
        self.calls += 1
        return self._d.decode(syndrome)[0]


def test_plugin_seam_runs_foreign_decoder_classes_per_shot():
    """sliding_window_circuit_mem with a decoder class of the caller's own, its own rate-keyword and method name (the reference's
    real plug point, doc/05 cells 10-11): window matrices from the C++ planner, one decoder object per window, one call per shot
    and window -- equal to the fixture the reference's own loop produced with the same inner decoder.  No GPU involved."""
    import quits_b200 as qb
    case = "bb72_r6_p3e-3_W5F3"
    g = decode_case(case)
    name = case_circuit(case)
    _, hz, lz = circuit_meta(name)
    n = 24
    params = dict(max_iter=10, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0", osd_order=0, precision="f64")
    d1, d2 = dict(params), dict(params)
    pred = qb.sliding_window_circuit_mem(g["det"][:n], qb.Circuit(circuit_text(name)), hz, lz, g["W"], g["F"], _ForeignDecoder,
                                         _ForeignDecoder, d1, d2, "my_rates", "my_rates", "run_it", "run_it")
    assert pred.dtype == np.int64 and np.array_equal(pred, g["pred_f64"][:n])
    assert "my_rates" in d1 and "my_rates" in d2           # the reference leaves the last priors in the caller's dicts


def test_count_logical_errors_equals_the_reference_idiom():
    """quits_b200.count_logical_errors == np.sum(np.any((obs - pred) % 2, axis=1)) (reference tests/test_sliding_window.py:83) for
    boolean flips against int64 predictions, odd K, negative / even integers, empty input, one thread or many."""
    import quits_b200 as qb
    rng = np.random.default_rng(3)
    for n, K in ((0, 12), (1, 1), (1000, 12), (70001, 7), (5000, 136)):
        obs = rng.random((n, K)) < 0.02
        pred = (obs ^ (rng.random((n, K)) < 0.003)).astype(np.int64)
        if n > 10:
            pred[3, 0] = -3
            pred[4, K - 1] = 2
        want = int(np.any((obs - pred) % 2, axis=1).sum())
        assert qb.count_logical_errors(obs, pred) == want
        assert qb.count_logical_errors(obs, pred, threads=1) == want
        assert qb.count_logical_errors(obs.astype(np.int64), pred.astype(np.bool_) if n == 0 else pred, threads=3) == want
    with pytest.raises(ValueError):
        qb.count_logical_errors(np.zeros((3, 2), bool), np.zeros((3, 3), np.int64))


def _load_bench():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("qb_bench_module", os.path.join(root, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_roofline_object_follows_the_committed_instruction_model():
    """bench.py's roofline: `achieved` = warp-instructions per edge-iteration (profiles/bp_inst_model.json, an ncu count) x the
    edge-iterations of the run / BP event time, against 148 x 4 x SM clock; the model is used only for the workload and decoder
    setting it was captured on, and the SURVEY 8(d) HBM figure stays beside it as hbm_model."""
    b = _load_bench()
    agg = {"bp_ms": 102.2 * 5, "bp_launches": 80, "bp_alg_bytes": 1.0e12, "bp_edge_iters": 2.69e11}
    clocks = {"sm_mhz": 1965}
    r = b.roofline("f64", agg, clocks, {"hbm_gbs": 6538.6})
    m = json.load(open(os.path.join(os.path.dirname(b.__file__), "profiles", "bp_inst_model.json")))["f64_minimum_sum_parallel"]
    assert r["bound"] == "issue" and r["kernel"] == "bp_kernel_ms2"
    assert abs(r["peak"] - 148 * 4 * 1.965) < 1e-6
    want = m["warp_inst_per_edge_iter"] * agg["bp_edge_iters"] / (agg["bp_ms"] / 1e3) / 1e9
    assert abs(r["achieved"] - want) < 1e-6 * want and abs(r["frac"] - want / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert r["hbm_model"]["bound"] == "hbm" and r["smem"]["frac"] < 1
    # another workload, or a decoder setting without a capture: no instruction model, no issue fraction
    b.WORKLOAD = "qlp1020_zxcol_r20_p5e-4"
    r = b.roofline("f64", agg, clocks, {"hbm_gbs": 6538.6})
    assert r["frac"] is None and r["achieved"] is None and "note" in r and r["traffic"] is None
    b.WORKLOAD = b.HEADLINE_WORKLOAD
    b.BP_KW["schedule"] = "serial"
    r = b.roofline("f64", agg, clocks, {"hbm_gbs": 6538.6})
    assert r["kernel"] == "bp_kernel_serial_slab" and r["frac"] is None


def test_bench_clock_sampler_degrades_without_a_gpu():
    b = _load_bench()
    s = b.ClockSampler(0)
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons", "samples", "source"} and out["reasons"] == []


def test_loaded_library_was_built_from_this_source_tree():
    """The shared library is git-ignored and travels prebuilt: its embedded source hash (qb_build_info) must equal the hash of the
    CUDA / C++ sources next to it, so a stale or foreign .so cannot pass for the committed code."""
    import quits_b200 as qbm
    from quits_b200 import build as qbuild
    info = qbm.build_info()
    assert "arch=sm_100a" in info and "fmad=off" in info
    assert ("src=" + qbuild.source_hash()) in info, info
