"""stim- / ldpc-shaped shims over the ORACLE, so the unmodified reference runs in the build container.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``install()`` registers fake ``stim`` and ``ldpc``
modules in ``sys.modules`` exposing exactly the surface the reference touches (SURVEY.md section 8b,
seam B4): ``stim.Circuit(text)`` (reference ``qldpc_code/bb.py:301``), ``len(c)`` / ``c[i].name`` /
``targets_copy()[j].qubit_value`` (``circuit.py:12-17``), ``compile_detector_sampler(seed=).sample``
(``simulation.py:23-27``), ``detector_error_model(decompose_errors=False)`` with ``flattened()`` /
``num_detectors`` / ``num_observables`` and instruction ``type`` / ``args_copy`` / ``targets_copy``
(``decoder/base.py:101-125,151``), and ``ldpc.bposd_decoder.BpOsdDecoder`` /
``ldpc.bplsd_decoder.BpLsdDecoder`` (``decoder/bposd.py:5``, ``decoder/bplsd.py:5``).

Used by ``tools/make_golden.py`` (run here, where /root/reference exists) to produce the fixtures in
``tests/golden/`` by driving the reference's own ``spacetime`` / ``sliding_window_circuit_mem``.
The product has its own, separate compat layer (``quits_b200/compat``) backed by the CUDA library.
"""
from __future__ import annotations

import sys
import types

import numpy as np

from . import stimtext


class _Target:
    def __init__(self, q):
        self.qubit_value = q
        self.value = q


class _CircuitInstr:
    def __init__(self, name, args, targets):
        self.name = name
        self._args = list(args)
        self._targets = list(targets)

    def targets_copy(self):
        return [_Target(t) for t in self._targets]

    def gate_args_copy(self):
        return list(self._args)


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


class _DemInstr:
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


class DetectorErrorModel:
    """Duck-typed stand-in for ``stim.DetectorErrorModel`` holding oracle.dem output."""

    def __init__(self, dem):
        self._dem = dem
        self.num_detectors = dem.n_det
        self.num_observables = dem.n_obs
        self.num_errors = len(dem.probs)

    def flattened(self):
        return self

    def __iter__(self):
        d = self._dem
        for p, dets, obs in zip(d.probs, d.dets, d.obs):
            yield _DemInstr("error", [float(p)], [DemTarget(x, False) for x in dets] + [DemTarget(x, True) for x in obs])

    def __len__(self):
        return self.num_errors


class _Sampler:
    def __init__(self, circuit, seed):
        self._c = circuit
        self._seed = seed

    def sample(self, shots, separate_observables=False, **kw):
        from . import cref
        if self._seed is None:
            seed = int(np.random.SeedSequence().entropy) & (2**63 - 1)
        else:
            seed = int(self._seed)
        det, obs = cref.sample(self._c.flat, seed, 0, int(shots))
        det = det.astype(np.bool_)
        obs = obs.astype(np.bool_)
        if separate_observables:
            return det, obs
        return det


class Circuit:
    """Text holder with the handful of ``stim.Circuit`` methods the reference uses."""

    def __init__(self, text=""):
        self.text = str(text)
        self._instrs = stimtext.parse(self.text)
        self._flat = None
        self._top = None

    @property
    def flat(self):
        if self._flat is None:
            self._flat = stimtext.flatten(self._instrs)
        return self._flat

    def _top_level(self):
        # stim fuses adjacent same-name/same-arg instructions; len()/[] index the fused list
        if self._top is None:
            top = []
            for ins in self._instrs:
                fus = ins.name not in stimtext.ANNOT and ins.name != "REPEAT"
                if fus and top and top[-1][3] and top[-1][0] == ins.name and top[-1][1] == ins.args:
                    top[-1][2].extend(ins.targets)
                else:
                    top.append([ins.name, ins.args, list(ins.targets), fus])
            self._top = top
        return self._top

    def __len__(self):
        return len(self._top_level())

    def __getitem__(self, i):
        name, args, targets, _ = self._top_level()[i]
        return _CircuitInstr(name, args, targets)

    def __str__(self):
        return stimtext.canonical_text(self.text)

    @property
    def num_detectors(self):
        return self.flat.n_det

    @property
    def num_observables(self):
        return self.flat.n_obs

    @property
    def num_qubits(self):
        return self.flat.n_qubits

    @property
    def num_measurements(self):
        return self.flat.n_meas

    def compile_detector_sampler(self, seed=None):
        return _Sampler(self, seed)

    def detector_error_model(self, decompose_errors=False, **kw):
        if decompose_errors:
            raise NotImplementedError("decompose_errors=True is not used on this path (decoder/base.py:151)")
        from . import dem as odem
        return DetectorErrorModel(odem.analyze(self.flat))


DEFAULT_PRECISION = "f64"      # tools/make_golden.py switches this to "f32" for the fixtures the GPU is held to


class _OracleBpDecoderBase:
    """ldpc-shaped per-shot decoder over the oracle's C BP(+OSD) (seam B3)."""
    _order_key = "osd_order"
    _method_key = "osd_method"

    def __init__(self, pcm, error_rate=None, error_channel=None, channel_probs=None, max_iter=0,
                 bp_method="minimum_sum", ms_scaling_factor=1.0, schedule="parallel", **kw):
        from . import cref
        import scipy.sparse as sp
        pcm = sp.csc_matrix(pcm)
        n = pcm.shape[1]
        if channel_probs is not None and error_channel is None:
            error_channel = channel_probs
        if error_channel is not None:
            priors = np.asarray(error_channel, dtype=np.float64)
        elif error_rate is not None:
            priors = np.full(n, float(error_rate))
        else:
            raise ValueError("error_rate / error_channel / channel_probs required")
        order = int(kw.get(self._order_key, 0))
        method = str(kw.get(self._method_key, getattr(self, "_default_method", "osd_0"))).lower().replace("_", "")
        if method.startswith("lsd"):
            # ldpc's lsd_method only matters for lsd_order > 0 (the order-0 solve is the same for lsd_0 / lsd_cs / lsd_e)
            if order != 0 and method != "lsd0":
                osd_method = "lsd_cs" if method == "lsdcs" else "lsd_e"         # per-cluster candidate sweep (cref.c, parity unpinned)
            else:
                osd_method, order = "lsd_0", 0
        else:
            osd_method = {"osd0": "osd_0", "osde": "osd_e", "exhaustive": "osd_e", "osdcs": "osd_cs", "combinationsweep": "osd_cs"}.get(method, "osd_0")
        self._dec = cref.BpOsd(pcm, priors, max_iter=max_iter if max_iter > 0 else n, bp_method=bp_method,
                               ms_scaling_factor=ms_scaling_factor, schedule=schedule,
                               precision=kw.get("precision", DEFAULT_PRECISION), osd_method=osd_method, osd_order=order)
        self.log_prob_ratios = None
        self.converge = None
        self.iter = None

    def decode(self, syndrome):
        e, llr, it, conv = self._dec.decode(np.asarray(syndrome) % 2)
        self.log_prob_ratios, self.iter, self.converge = llr, it, conv
        return e


class BpOsdDecoder(_OracleBpDecoderBase):
    pass


class BpLsdDecoder(_OracleBpDecoderBase):
    _order_key = "lsd_order"
    _method_key = "lsd_method"
    _default_method = "lsd_0"


def install(force=False):
    """Register the shims as ``stim`` / ``ldpc`` unless the real packages are importable."""
    if not force:
        try:
            import stim as _real_stim  # noqa: F401
            import ldpc as _real_ldpc  # noqa: F401
            return False
        except ImportError:
            pass
    stim = types.ModuleType("stim")
    stim.Circuit = Circuit
    stim.DetectorErrorModel = DetectorErrorModel
    stim.DemTarget = DemTarget
    stim.__oracle_shim__ = True
    ldpc = types.ModuleType("ldpc")
    bposd = types.ModuleType("ldpc.bposd_decoder")
    bposd.BpOsdDecoder = BpOsdDecoder
    bplsd = types.ModuleType("ldpc.bplsd_decoder")
    bplsd.BpLsdDecoder = BpLsdDecoder
    ldpc.bposd_decoder = bposd
    ldpc.bplsd_decoder = bplsd
    ldpc.BpOsdDecoder = BpOsdDecoder
    ldpc.__oracle_shim__ = True
    sys.modules["stim"] = stim
    sys.modules["ldpc"] = ldpc
    sys.modules["ldpc.bposd_decoder"] = bposd
    sys.modules["ldpc.bplsd_decoder"] = bplsd
    return True
