"""Run the UNMODIFIED reference package on the engine.

QUITS reaches its arithmetic through two wheels, at a handful of names (seam B4): ``stim.Circuit(text)``
(``qldpc_code/bb.py:301``, ``circuit_construction/cardinal.py:267``, ``zxcoloration.py:270``), ``len(c)`` / ``c[i].name`` /
``c[i].targets_copy()[j].qubit_value`` (``circuit.py:12-17``), ``c.compile_detector_sampler(seed=).sample(shots=,
separate_observables=True)`` (``simulation.py:23-27``), ``c.detector_error_model(decompose_errors=False)`` and the flattened DEM
iteration (``decoder/base.py:101-125,151``), ``ldpc.bposd_decoder.BpOsdDecoder`` / ``ldpc.bplsd_decoder.BpLsdDecoder``
(``decoder/bposd.py:5``, ``decoder/bplsd.py:5``).  ``install()`` registers engine-backed modules under those names, so that
``import quits`` -- code families, circuit builders, ``get_stim_mem_result``, ``spacetime``, the sliding-window functions --
runs as it is, with the frame kernel behind the sampler and the BP / OSD / LSD kernels behind the inner decoders.

    import quits_b200.compat as compat
    compat.install()            # before ``import quits``; a no-op when the real stim and ldpc are importable (force=True overrides)
    import quits

The reference's own window loop then calls ``decoder.decode(syndrome)`` once per shot and window (seam B3): correct, but one
kernel launch per call.  The batched path is ``quits_b200.sliding_window_*`` (same signatures), which recognises these
classes and decodes all shots at once.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from .circuit import Circuit, DemTarget, DetectorErrorModel
from .decoder.inner import BpLsdDecoder, BpOsdDecoder

_GATES_WITH_QUBIT_TARGETS = {"R", "RX", "H", "CX", "CNOT", "M", "MX", "MR", "X_ERROR", "Z_ERROR", "DEPOLARIZE1", "DEPOLARIZE2",
                             "PAULI_CHANNEL_1", "PAULI_CHANNEL_2", "TICK"}


class _GateTarget:
    def __init__(self, q):
        self.qubit_value = q
        self.value = q

    def __repr__(self):
        return str(self.qubit_value)


class _Instruction:
    """One top-level instruction of the text: the view ``circuit.py:12-17`` of the reference walks."""

    def __init__(self, name, args, targets):
        self.name = name
        self._args = args
        self._targets = targets

    def targets_copy(self):
        return list(self._targets)

    def gate_args_copy(self):
        return list(self._args)

    def __repr__(self):
        return "%s %s" % (self.name, " ".join(map(repr, self._targets)))


def _top_level_instructions(text):
    out, depth = [], 0
    for raw in text.split("\n"):
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("}"):
            depth -= 1
            continue
        opens = line.endswith("{")
        if depth == 0:
            head = line[:-1].strip() if opens else line
            tok = head.split()
            name, args = tok[0], []
            if "(" in name:
                name, rest = name.split("(", 1)
                args = [float(x) for x in rest.rstrip(")").split(",") if x]
            name = name.upper()
            targets = []
            if name in _GATES_WITH_QUBIT_TARGETS:
                targets = [_GateTarget(int(t)) for t in tok[1:] if t.lstrip("!").isdigit()]
            out.append(_Instruction(name, args, targets))
        if opens:
            depth += 1
    return out


class _Sampler:
    def __init__(self, circuit, seed):
        self._c = circuit
        self._seed = seed

    def sample(self, shots, separate_observables=False, **kw):
        seed = int(np.random.SeedSequence().entropy) & (2 ** 63 - 1) if self._seed is None else int(self._seed)
        det, obs = self._c.sample(int(shots), seed)
        return (det, obs) if separate_observables else det


class StimCircuit(Circuit):
    """``stim.Circuit``-shaped: the engine's circuit (C++ front end, frame kernel, DEM analyser) plus the few methods the
    reference calls on a ``stim.Circuit``."""

    def __init__(self, text=""):
        super().__init__(str(text))
        self._top = None

    def _instructions(self):
        if self._top is None:
            self._top = _top_level_instructions(self.text)
        return self._top

    def __len__(self):
        return len(self._instructions())

    def __getitem__(self, i):
        return self._instructions()[i]

    def compile_detector_sampler(self, seed=None, **kw):
        return _Sampler(self, seed)


def install(force: bool = False) -> bool:
    """Register the engine under ``stim`` / ``ldpc`` / ``ldpc.bposd_decoder`` / ``ldpc.bplsd_decoder``.  Returns False (and does
    nothing) when the real packages are importable and ``force`` is not set."""
    if not force:
        try:
            import ldpc as _real_ldpc  # noqa: F401
            import stim as _real_stim  # noqa: F401
            if not getattr(_real_stim, "__quits_b200__", False) and not getattr(_real_stim, "__oracle_shim__", False):
                return False
        except ImportError:
            pass
    stim = types.ModuleType("stim")
    stim.Circuit = StimCircuit
    stim.DetectorErrorModel = DetectorErrorModel
    stim.DemTarget = DemTarget
    stim.__quits_b200__ = True
    ldpc = types.ModuleType("ldpc")
    bposd = types.ModuleType("ldpc.bposd_decoder")
    bposd.BpOsdDecoder = BpOsdDecoder
    bplsd = types.ModuleType("ldpc.bplsd_decoder")
    bplsd.BpLsdDecoder = BpLsdDecoder
    ldpc.bposd_decoder, ldpc.bplsd_decoder = bposd, bplsd
    ldpc.BpOsdDecoder, ldpc.BpLsdDecoder = BpOsdDecoder, BpLsdDecoder
    ldpc.__quits_b200__ = True
    sys.modules["stim"] = stim
    sys.modules["ldpc"] = ldpc
    sys.modules["ldpc.bposd_decoder"] = bposd
    sys.modules["ldpc.bplsd_decoder"] = bplsd
    return True


__all__ = ["StimCircuit", "install"]
