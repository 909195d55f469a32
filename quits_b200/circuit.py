"""Host objects over the C ABI: device context, parsed circuit, detector error model.

``Circuit`` stands where ``stim.Circuit`` stands in the reference (built from the Stim text the reference's
builders emit, reference ``src/quits/qldpc_code/bb.py:301``); ``DetectorErrorModel`` stands where
``circuit.detector_error_model(decompose_errors=False)`` stands (``decoder/base.py:151``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _native as N

OP_NAMES = ["R", "RX", "H", "CX", "M", "MX", "MR", "X_ERROR", "Z_ERROR", "DEPOLARIZE1", "DEPOLARIZE2", "DETECTOR",
            "OBSERVABLE_INCLUDE"]

_ctx_lock = threading.Lock()
_contexts = {}


class Context:
    """One CUDA device + stream.  ``Context.default()`` picks LOCAL_RANK (one process per GPU) or device 0."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        N.check(N.lib().qb_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @classmethod
    def default(cls, device=None) -> "Context":
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            n = N.device_count()
            if n > 0:
                device %= n
        with _ctx_lock:
            ctx = _contexts.get(device)
            if ctx is None:
                ctx = _contexts[device] = cls(device)
            return ctx

    def synchronize(self):
        N.check(N.lib().qb_ctx_synchronize(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and N.alive():
                N.lib().qb_ctx_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _default_ctx() -> "Context":
    """The context calls without an explicit one run on: slot 0 of quits_b200.devices."""
    from . import devices as dv
    return dv.slot_context(0)


def circuit_text(circuit) -> str:
    """Stim text of whatever the caller holds: our Circuit, a str, or a stim.Circuit (``str()`` prints Stim text)."""
    if isinstance(circuit, Circuit):
        return circuit.text
    if isinstance(circuit, (bytes, bytearray)):
        return bytes(circuit).decode()
    if isinstance(circuit, str):
        return circuit
    text = getattr(circuit, "text", None)
    if isinstance(text, str):
        return text
    return str(circuit)


class Circuit:
    """Parsed Stim-text circuit (host side: flattened op list + device tape)."""

    def __init__(self, text=""):
        self.text = circuit_text(text) if not isinstance(text, str) else text
        raw = self.text.encode()
        h = C.c_void_p()
        N.check(N.lib().qb_circuit_parse(raw, len(raw), C.byref(h)))
        self._h = h
        info = N.QbCircuitInfo()
        N.check(N.lib().qb_circuit_get_info(self._h, C.byref(info)))
        self.num_qubits = info.n_qubits
        self.num_measurements = info.n_measurements
        self.num_detectors = info.n_detectors
        self.num_observables = info.n_observables
        self.num_flat_ops = info.n_flat_ops
        self.num_tape_ops = info.n_tape_ops
        self.num_noise_sites = info.n_noise_sites
        self.ring = info.ring
        self._dem = None
        self._clones = {}         # slot -> Circuit: the device tape of a parsed circuit belongs to one device at a time

    @classmethod
    def of(cls, circuit) -> "Circuit":
        return circuit if isinstance(circuit, cls) else cls(circuit_text(circuit))

    def __del__(self):
        try:
            if getattr(self, "_h", None) and N.alive():
                N.lib().qb_circuit_free(self._h)
                self._h = None
        except Exception:
            pass

    def __str__(self):
        return self.text

    def flat(self):
        """(kind int32[n], arg float64[n], tstart int64[n+1], targets int32[...]) of the flattened op list."""
        n = self.num_flat_ops
        kind = np.zeros(n, dtype=np.int32)
        arg = np.zeros(n, dtype=np.float64)
        tstart = np.zeros(n + 1, dtype=np.int64)
        nt = C.c_int64(0)
        N.check(N.lib().qb_circuit_flat(self._h, N.ptr(kind), N.ptr(arg), N.ptr(tstart), None, C.byref(nt)))
        targets = np.zeros(max(nt.value, 1), dtype=np.int32)
        N.check(N.lib().qb_circuit_flat(self._h, None, None, None, N.ptr(targets), None))
        return kind, arg, tstart, targets[:nt.value]

    def detector_error_model(self, decompose_errors=False, **kw) -> "DetectorErrorModel":
        if decompose_errors:
            raise NotImplementedError("decompose_errors=True is not used on this path (reference decoder/base.py:151)")
        if self._dem is None:
            self._dem = DetectorErrorModel(self)
        return self._dem

    def for_slot(self, slot: int) -> "Circuit":
        """This circuit for device slot ``slot`` (quits_b200.devices): slot 0 is the object itself, later slots get a parsed copy."""
        if slot == 0:
            return self
        c = self._clones.get(slot)
        if c is None:
            c = self._clones[slot] = Circuit(self.text)
        return c

    # ---- sampling (reference simulation.py:22-27)
    def sample(self, shots, seed, shot0=0, ctx=None, packed=False):
        """Shots [shot0, shot0 + shots) of ``seed``.  Without an explicit context the shots are spread over every device of
        quits_b200.devices (contiguous ranges; shot s is the same bits wherever it is sampled)."""
        from . import devices as dv
        shots = int(shots)
        D, K = self.num_detectors, self.num_observables
        if packed:
            det = N.empty((shots, max(1, (D + 63) // 64)), np.uint64)
            obs = N.empty((shots, max(1, (K + 63) // 64)), np.uint64)
            fn = N.lib().qb_sample_packed
        else:
            det = N.empty((shots, D), np.bool_)
            obs = N.empty((shots, K), np.bool_)
            fn = N.lib().qb_sample
        if ctx is not None or int(shot0) % 64:
            N.check(fn((ctx or _default_ctx())._h, self._h, int(seed), int(shot0), shots, N.ptr(det), N.ptr(obs)))
            return det, obs

        def part(slot, lo, hi):
            c = self.for_slot(slot)
            N.check(fn(dv.slot_context(slot)._h, c._h, int(seed), int(shot0) + lo, hi - lo, N.ptr(det[lo:hi]), N.ptr(obs[lo:hi])))

        dv.run_split(shots, part)
        return det, obs

    def inject(self, op_idx, tgt_idx, codes, shots=None, n_shots=None, ctx=None):
        """Explicit-fault propagation: fault f goes into shot ``shots[f]`` (default: one fault per shot)."""
        ctx = ctx or _default_ctx()
        op_idx = np.ascontiguousarray(op_idx, dtype=np.int32)
        tgt_idx = np.ascontiguousarray(tgt_idx, dtype=np.int32)
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        nf = len(op_idx)
        shots = np.arange(nf, dtype=np.int64) if shots is None else np.ascontiguousarray(shots, dtype=np.int64)
        if n_shots is None:
            n_shots = int(shots.max()) + 1 if nf else 0
        det = np.zeros((n_shots, self.num_detectors), dtype=np.bool_)
        obs = np.zeros((n_shots, self.num_observables), dtype=np.bool_)
        N.check(N.lib().qb_sample_faults(ctx._h, self._h, nf, N.ptr(op_idx), N.ptr(tgt_idx), N.ptr(codes), N.ptr(shots),
                                         int(n_shots), N.ptr(det), N.ptr(obs)))
        return det, obs


class DetectorErrorModel:
    """Stim-ordered error list of a circuit plus the QUITS check-matrix view of it."""

    def __init__(self, circuit: Circuit):
        h = C.c_void_p()
        N.check(N.lib().qb_dem_from_circuit(circuit._h, C.byref(h)))
        self._h = h
        sizes = np.zeros(9, dtype=np.int64)
        N.check(N.lib().qb_dem_sizes(self._h, N.ptr(sizes)))
        (self.num_detectors, self.num_observables, self.num_errors, self._nnz_det, self._nnz_obs, self.num_columns,
         self._nnz_h, self._nnz_l, self.num_detectorless) = (int(x) for x in sizes)

    def __del__(self):
        try:
            if getattr(self, "_h", None) and N.alive():
                N.lib().qb_dem_free(self._h)
                self._h = None
        except Exception:
            pass

    def errors(self):
        """dict with probs, det_ptr/det_idx, obs_ptr/obs_idx (CSR over errors, stim order) and representative faults."""
        n = self.num_errors
        out = {"probs": np.zeros(n, dtype=np.float64), "det_ptr": np.zeros(n + 1, dtype=np.int64),
               "det_idx": np.zeros(max(self._nnz_det, 1), dtype=np.int32), "obs_ptr": np.zeros(n + 1, dtype=np.int64),
               "obs_idx": np.zeros(max(self._nnz_obs, 1), dtype=np.int32), "rep_op": np.zeros(n, dtype=np.int32),
               "rep_tgt": np.zeros(n, dtype=np.int32), "rep_code": np.zeros(n, dtype=np.int32)}
        N.check(N.lib().qb_dem_errors(self._h, *[N.ptr(out[k]) for k in ("probs", "det_ptr", "det_idx", "obs_ptr", "obs_idx",
                                                                           "rep_op", "rep_tgt", "rep_code")]))
        out["det_idx"] = out["det_idx"][:self._nnz_det]
        out["obs_idx"] = out["obs_idx"][:self._nnz_obs]
        return out

    def matrix_arrays(self):
        c = self.num_columns
        h_ptr = np.zeros(c + 1, dtype=np.int64)
        h_idx = np.zeros(max(self._nnz_h, 1), dtype=np.int32)
        l_ptr = np.zeros(c + 1, dtype=np.int64)
        l_idx = np.zeros(max(self._nnz_l, 1), dtype=np.int32)
        priors = np.zeros(c, dtype=np.float64)
        N.check(N.lib().qb_dem_matrix(self._h, N.ptr(h_ptr), N.ptr(h_idx), N.ptr(l_ptr), N.ptr(l_idx), N.ptr(priors)))
        return h_ptr, h_idx[:self._nnz_h], l_ptr, l_idx[:self._nnz_l], priors

    # ---- the slice of stim.DetectorErrorModel the reference iterates over (decoder/base.py:101-125)
    def flattened(self):
        return self

    def __len__(self):
        return self.num_errors

    def __iter__(self):
        e = self.errors()
        for i in range(self.num_errors):
            d = e["det_idx"][e["det_ptr"][i]:e["det_ptr"][i + 1]]
            o = e["obs_idx"][e["obs_ptr"][i]:e["obs_ptr"][i + 1]]
            yield DemInstruction("error", [float(e["probs"][i])], [DemTarget(int(x), False) for x in d] + [DemTarget(int(x), True) for x in o])


class DemTarget:
    def __init__(self, val, is_obs):
        self.val = int(val)
        self._obs = bool(is_obs)

    def is_relative_detector_id(self):
        return not self._obs

    def is_logical_observable_id(self):
        return self._obs

    def is_separator(self):
        return False

    def __repr__(self):
        return ("L%d" if self._obs else "D%d") % self.val


class DemInstruction:
    def __init__(self, type_, args, targets):
        self.type = type_
        self._args = args
        self._targets = targets

    def args_copy(self):
        return list(self._args)

    def targets_copy(self):
        return list(self._targets)

    def __repr__(self):
        return "%s(%r) %s" % (self.type, self._args[0] if self._args else "", " ".join(map(repr, self._targets)))
