"""ctypes front-end of oracle/cref.c (the C restatement of the hot path).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .stimtext import FlatCircuit

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqoracle.so")
_SRC = [os.path.join(_HERE, "cref.c"), os.path.join(_HERE, "bp_impl.inc")]
_lib = None

OPK = {"R": 0, "RX": 1, "H": 2, "CX": 3, "M": 4, "MX": 5, "MR": 6, "X_ERROR": 7, "Z_ERROR": 8, "DEPOLARIZE1": 9,
       "DEPOLARIZE2": 10, "DETECTOR": 11, "OBSERVABLE_INCLUDE": 12}


def build(force: bool = False) -> str:
    """gcc -O3 -fopenmp the C oracle into oracle/libqoracle.so.  -march=x86-64-v3 (AVX2 / BMI2) rather than -march=native: the
    library is built in the build container and travels to the GPU box, whose host CPU may be another model; no FP contraction
    (the GPU kernels are held bit-exact to this arithmetic)."""
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in _SRC):
        return _SO
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-mtune=generic", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-Wno-stringop-overflow", "-shared", "-fPIC", "-o", _SO, _SRC[0], "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.qo_bp_create.restype = C.c_void_p
        _lib.qo_bp_free.argtypes = [C.c_void_p]
        _lib.qo_num_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def flat_arrays(fc: FlatCircuit):
    kind = np.array([OPK[o.name] for o in fc.ops], dtype=np.int32)
    arg = np.array([o.arg for o in fc.ops], dtype=np.float64)
    tstart = np.zeros(len(fc.ops) + 1, dtype=np.int64)
    tstart[1:] = np.cumsum([len(o.targets) for o in fc.ops])
    targets = np.array([t for o in fc.ops for t in o.targets], dtype=np.int32)
    if targets.size == 0:
        targets = np.zeros(1, dtype=np.int32)
    return kind, arg, tstart, targets


def philox(k0, k1, ctr):
    ctr = np.asarray(ctr, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().qo_philox(C.c_uint32(k0), C.c_uint32(k1), _p(ctr, C.c_uint32), _p(out, C.c_uint32))
    return out


def noise_tables(p: float):
    t1 = C.c_uint32(0)
    c = np.zeros(64, dtype=np.uint64)
    lib().qo_noise_tables(C.c_double(p), C.byref(t1), _p(c, C.c_uint64))
    return int(t1.value), c


def _run(fc, seed, word0, nwords, inj=None, nthreads=0):
    kind, arg, tstart, targets = flat_arrays(fc)
    det = np.zeros((max(fc.n_det, 1), nwords), dtype=np.uint64)
    obs = np.zeros((max(fc.n_obs, 1), nwords), dtype=np.uint64)
    if inj is None:
        n_inj = -1
        i_op = i_t = i_c = np.zeros(1, dtype=np.int32)
        i_s = np.zeros(1, dtype=np.int64)
    else:
        i_op, i_t, i_c, i_s = (np.ascontiguousarray(inj[0], dtype=np.int32), np.ascontiguousarray(inj[1], dtype=np.int32),
                               np.ascontiguousarray(inj[2], dtype=np.int32), np.ascontiguousarray(inj[3], dtype=np.int64))
        order = np.argsort(i_op, kind="stable")
        i_op, i_t, i_c, i_s = i_op[order].copy(), i_t[order].copy(), i_c[order].copy(), i_s[order].copy()
        n_inj = len(i_op)
    rc = lib().qo_run(C.c_int(len(kind)), _p(kind, C.c_int32), _p(arg, C.c_double), _p(tstart, C.c_int64),
                      _p(targets, C.c_int32), C.c_int(fc.n_qubits), C.c_int(fc.n_meas), C.c_int(fc.n_det), C.c_int(fc.n_obs),
                      C.c_uint64(seed), C.c_uint64(word0), C.c_uint64(nwords), _p(det, C.c_uint64), _p(obs, C.c_uint64),
                      C.c_int64(n_inj), _p(i_op, C.c_int32), _p(i_t, C.c_int32), _p(i_c, C.c_int32), _p(i_s, C.c_int64),
                      C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("qo_run failed with code %d" % rc)
    return det[:fc.n_det], obs[:fc.n_obs]


def unpack_planes(planes: np.ndarray, nshots: int) -> np.ndarray:
    """[bits][nwords] u64 bit-sliced -> uint8 [nshots][bits]."""
    b = np.unpackbits(planes.view(np.uint8), axis=1, bitorder="little")     # [bits][nwords*64]
    return np.ascontiguousarray(b[:, :nshots].T)


def sample_planes(fc, seed, word0, nwords, nthreads=0):
    return _run(fc, seed, word0, nwords, None, nthreads)


def sample(fc, seed, shot0, nshots, nthreads=0):
    """uint8 [nshots][D], [nshots][K] for global shots shot0 .. shot0+nshots (shot0 must be a multiple of 64)."""
    assert shot0 % 64 == 0
    nwords = (nshots + 63) // 64
    det, obs = _run(fc, seed, shot0 // 64, nwords, None, nthreads)
    return unpack_planes(det, nshots), unpack_planes(obs, nshots)


def inject(fc, op_idx, tgt_idx, codes):
    """One explicit fault per shot (shot f gets fault f); noise instructions are otherwise skipped."""
    n = len(op_idx)
    nwords = (n + 63) // 64
    det, obs = _run(fc, 0, 0, nwords, (op_idx, tgt_idx, codes, np.arange(n)))
    return unpack_planes(det, n), unpack_planes(obs, n)


class BpOsd:
    """One window's decoder: ldpc.BpOsdDecoder-shaped (decode one syndrome at a time)."""
    METHODS = {"minimum_sum": 0, "min_sum": 0, "ms": 0, "msl": 0, "product_sum": 1, "ps": 1, "psl": 1, 0: 0, 1: 1}
    SCHEDULES = {"parallel": 0, "serial": 1, 0: 0, 1: 1}

    OSD_METHODS = {"osd_0": 0, "osd0": 0, "osd_e": 1, "osde": 1, "exhaustive": 1, "osd_cs": 2, "osdcs": 2, "combination_sweep": 2,
                   "lsd_0": 3, "lsd0": 3, "lsd_e": 4, "lsd_cs": 5, 0: 0, 1: 1, 2: 2, 3: 3, 4: 4, 5: 5}

    def __init__(self, pcm, priors, max_iter, bp_method="minimum_sum", ms_scaling_factor=1.0, schedule="parallel",
                 precision="f64", osd=True, osd_method="osd_0", osd_order=0):
        import scipy.sparse as sp
        pcm = sp.csc_matrix(pcm)
        pcm.sort_indices()
        self.m, self.n = pcm.shape
        self._indptr = np.ascontiguousarray(pcm.indptr, dtype=np.int64)
        self._indices = np.ascontiguousarray(pcm.indices, dtype=np.int32)
        self._priors = np.ascontiguousarray(priors, dtype=np.float64)
        assert len(self._priors) == self.n
        self.h = lib().qo_bp_create(C.c_int(self.m), C.c_int(self.n), _p(self._indptr, C.c_int64), _p(self._indices, C.c_int32),
                                    _p(self._priors, C.c_double), C.c_int(int(max_iter)), C.c_int(self.METHODS[bp_method]),
                                    C.c_int(self.SCHEDULES[schedule]), C.c_double(float(ms_scaling_factor)),
                                    C.c_int(32 if precision in ("f32", 32) else 64), C.c_int(1 if osd else 0))
        lib().qo_bp_set_osd(C.c_void_p(self.h), C.c_int(self.OSD_METHODS[str(osd_method).lower() if isinstance(osd_method, str) else osd_method]),
                            C.c_int(int(osd_order)))
        self.used_osd = 0

    def __del__(self):
        try:
            if self.h:
                lib().qo_bp_free(C.c_void_p(self.h))
                self.h = None
        except Exception:
            pass

    def decode(self, syndrome):
        s = np.ascontiguousarray(np.asarray(syndrome) % 2, dtype=np.uint8)
        assert s.shape == (self.m,)
        e = np.zeros(self.n, dtype=np.uint8)
        llr = np.zeros(self.n, dtype=np.float64)
        it = C.c_int(0)
        used = C.c_int(0)
        conv = lib().qo_bp_decode(C.c_void_p(self.h), _p(s, C.c_uint8), _p(e, C.c_uint8), _p(llr, C.c_double), C.byref(it),
                                  C.byref(used))
        self.used_osd = used.value
        return e, llr, it.value, bool(conv)

    @staticmethod
    def lsd_diag():
        """(row operations created, bits added, merges) of this thread's last LSD call."""
        out = np.zeros(3, dtype=np.int64)
        lib().qo_lsd_diag(_p(out, C.c_int64))
        return tuple(int(x) for x in out)


def sw_decode(windows, m, K, det, nthreads=0, **bpkw):
    """Restated sliding-window loop over prepared windows.

    windows: list of dicts {H (csc), priors, L (csc K x ncommit), U (csc m x ncommit or None), row0}.
    det: uint8 [N][D].  Returns (pred uint8 [N][K], stats int64 [n_windows][3]).
    """
    import scipy.sparse as sp
    det = np.ascontiguousarray(det, dtype=np.uint8)
    N, D = det.shape
    nw = len(windows)
    decs = [BpOsd(w["H"], w["priors"], **bpkw) for w in windows]
    hs = (C.c_void_p * nw)(*[d.h for d in decs])
    row0 = np.array([w["row0"] for w in windows], dtype=np.int32)
    ncommit = np.array([w["L"].shape[1] for w in windows], dtype=np.int32)
    keep = []
    Lp, Li, Up, Ui = [], [], [], []
    for w in windows:
        L = sp.csc_matrix(w["L"]); L.sort_indices()
        lp = np.ascontiguousarray(L.indptr, dtype=np.int64); li = np.ascontiguousarray(L.indices, dtype=np.int32)
        if li.size == 0:
            li = np.zeros(1, dtype=np.int32)
        keep += [lp, li]
        Lp.append(lp.ctypes.data); Li.append(li.ctypes.data)
        if w.get("U") is not None:
            U = sp.csc_matrix(w["U"]); U.sort_indices()
            up = np.ascontiguousarray(U.indptr, dtype=np.int64); ui = np.ascontiguousarray(U.indices, dtype=np.int32)
            if ui.size == 0:
                ui = np.zeros(1, dtype=np.int32)
            keep += [up, ui]
            Up.append(up.ctypes.data); Ui.append(ui.ctypes.data)
        else:
            Up.append(None); Ui.append(None)
    arr = lambda xs: (C.c_void_p * nw)(*xs)
    pred = np.zeros((N, K), dtype=np.uint8)
    stats = np.zeros((nw, 3), dtype=np.int64)
    rc = lib().qo_sw_decode(C.c_int(nw), hs, _p(row0, C.c_int32), _p(ncommit, C.c_int32), arr(Lp), arr(Li), arr(Up), arr(Ui),
                            C.c_int(m), C.c_int(K), C.c_int(D), _p(det, C.c_uint8), C.c_int64(N), _p(pred, C.c_uint8),
                            _p(stats, C.c_int64), C.c_int(nthreads))
    assert rc == 0
    del keep
    return pred, stats


def num_threads() -> int:
    return int(lib().qo_num_threads())
