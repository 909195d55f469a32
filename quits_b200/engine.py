"""Device-side objects: the batched sliding-window decoder and the fused Monte-Carlo runner.

``SlidingWindowDecoder`` is the engine behind the drop-in ``sliding_window_*_circuit_mem`` functions
(reference ``src/quits/decoder/sliding_window.py:104-188``); ``MonteCarlo`` chains sampling, decoding and the
logical-error count on the device (``get_stim_mem_result`` -> ``sliding_window_bposd_circuit_mem`` -> the caller's
``pL`` reduction, reference ``tests/test_sliding_window.py:66-84``) and shards shots across GPUs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N
from .circuit import Circuit, Context, _default_ctx
from .decoder.base import WindowPlan

_BP_METHODS = {"minimum_sum": 0, "min_sum": 0, "ms": 0, "msl": 0, "product_sum": 1, "prod_sum": 1, "ps": 1, "psl": 1}
_SCHEDULES = {"parallel": 0, "p": 0, "serial": 1, "s": 1}
_OSD_METHODS = {"osd_0": 0, "osd0": 0, "osd_e": 1, "osde": 1, "exhaustive": 1, "osd_cs": 2, "osdcs": 2, "combination_sweep": 2,
                "lsd_0": 3, "lsd0": 3, "lsd_e": 4, "lsde": 4, "lsd_cs": 5, "lsdcs": 5, "off": -1, "none": -1}


def bp_options(bp_method="minimum_sum", max_iter=0, schedule="parallel", osd_method="osd_0", osd_order=0, ms_scaling_factor=1.0,
               precision="f64", capacity=0, profile=False, lanes=0, **unknown) -> N.QbBpOpts:
    """Translate ldpc.BpOsdDecoder-style kwargs (reference decoder/bposd.py:74-83) into the C option block."""
    for k, v in unknown.items():
        if k not in ("channel_probs", "error_rate", "error_channel", "input_vector_type", "omp_thread_count",
                     "random_schedule_seed", "serial_schedule_order", "osd_on"):
            raise TypeError("unexpected decoder option %r" % k)
        # options ldpc accepts that would change what is decoded: only their defaults are implemented (no silent ignoring)
        if k == "serial_schedule_order" and v is not None:
            raise NotImplementedError("serial_schedule_order: only the default order (column index) runs on the GPU path")
        if k == "random_schedule_seed" and v not in (None, 0, -1):
            raise NotImplementedError("random_schedule_seed: randomised serial schedules are not implemented on the GPU path")
        if k == "input_vector_type" and str(v).lower() not in ("syndrome", "auto", "0", "none"):
            raise NotImplementedError("input_vector_type=%r: only syndrome input is implemented on the GPU path" % (v,))
        if k == "osd_on" and not v:
            osd_method = "off"
    # ldpc also accepts small integers here, with its own numbering (bp_method 0 = product_sum there); an integer passed on from
    # ldpc-style code would silently select another algorithm through the engine's enums, so only names are taken
    for name, v in (("bp_method", bp_method), ("schedule", schedule), ("osd_method", osd_method)):
        if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
            raise ValueError("%s must be given by name (e.g. 'minimum_sum', 'parallel', 'osd_cs'), got the integer %r" % (name, v))
    o = N.QbBpOpts()
    try:
        o.bp_method = _BP_METHODS[str(bp_method).lower()]
        o.schedule = _SCHEDULES[str(schedule).lower()]
        o.osd_method = _OSD_METHODS[str(osd_method).lower()]
    except KeyError as e:
        raise ValueError("unknown decoder option value %s" % e) from None
    if precision not in ("f64", "f32", 64, 32):
        raise ValueError("precision must be 'f64' (what ldpc computes in; default) or 'f32'")
    o.precision = 32 if precision in ("f32", 32) else 64
    o.max_iter = int(max_iter)
    o.ms_scaling_factor = float(ms_scaling_factor)
    o.osd_order = int(osd_order)
    o.capacity = int(capacity)
    o.profile = 1 if profile else 0
    o.lanes = int(lanes)
    return o


class SlidingWindowDecoder:
    """All windows of one circuit resident on a GPU; decodes batches of shots (K3/K4 kernels).

    Built without an explicit context the decoder belongs to the default device and, for large calls, spreads the shots over
    every device of ``quits_b200.devices``: each further device gets its own copy of the device tables on first use (the window
    plan is host data and is shared)."""

    def __init__(self, circuit, m: int, W: int, F: int, num_cor_rounds: int = -1, ctx: Context = None, **bp_kwargs):
        self._fanout = ctx is None
        self.ctx = ctx or _default_ctx()
        self.circuit = Circuit.of(circuit)
        self.plan = WindowPlan(self.circuit.detector_error_model(), m, W, F, num_cor_rounds)
        self._bp_kwargs = dict(bp_kwargs)
        self.opts = bp_options(**bp_kwargs)
        h = C.c_void_p()
        N.check(N.lib().qb_sw_create(self.ctx._h, self.plan._h, C.byref(self.opts), C.byref(h)))
        self._h = h
        self.K, self.D = self.plan.K, self.plan.D
        self.stats = N.QbStats()
        self._peers = {}

    @classmethod
    def from_plan(cls, plan: WindowPlan, ctx: Context = None, **bp_kwargs) -> "SlidingWindowDecoder":
        """Decoder over an explicit window plan (``WindowPlan.explicit``), e.g. the phenomenological windows."""
        self = cls.__new__(cls)
        self._fanout = ctx is None
        self.ctx = ctx or _default_ctx()
        self.circuit = None
        self.plan = plan
        self._bp_kwargs = dict(bp_kwargs)
        self._peers = {}
        self.opts = bp_options(**bp_kwargs)
        h = C.c_void_p()
        N.check(N.lib().qb_sw_create(self.ctx._h, plan._h, C.byref(self.opts), C.byref(h)))
        self._h = h
        self.K, self.D = plan.K, plan.D
        self.stats = N.QbStats()
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None) and N.alive():
                N.lib().qb_sw_free(self._h)
                self._h = None
        except Exception:
            pass

    def _peer(self, slot: int) -> "SlidingWindowDecoder":
        """The same decoder on the device of slot ``slot`` (slot 0: this object)."""
        if slot == 0:
            return self
        from . import devices as dv
        d = self._peers.get(slot)
        if d is None:
            d = self._peers[slot] = SlidingWindowDecoder.from_plan(self.plan, ctx=dv.slot_context(slot), **self._bp_kwargs)
        return d

    def _split(self, n, call):
        """Run ``call(decoder, lo, hi)`` over the shot ranges of the device split (one range on this decoder when it was built on
        an explicit context)."""
        if not self._fanout:
            return call(self, 0, n)
        from . import devices as dv
        parts = dv.plan_split(n)
        for slot, _, _ in parts:                       # device tables are set up on the calling thread, one device after the other
            self._peer(slot)
        dv.run_split(n, lambda slot, lo, hi: call(self._peer(slot), lo, hi))

    def decode(self, det) -> np.ndarray:
        """det: [N, D] array of 0/1 (bool or integer) on the host -> int64 [N, K] predictions."""
        det = np.asarray(det)
        if det.ndim != 2 or det.shape[1] != self.D:
            raise ValueError("expected detection events of shape (N, %d), got %r" % (self.D, det.shape))
        if det.dtype != np.bool_ and det.dtype != np.uint8:
            det = (det % 2).astype(np.uint8)
        det = np.ascontiguousarray(det).view(np.uint8)
        pred = N.empty((det.shape[0], self.K), np.int64)
        if self.K == 0 or det.shape[0] == 0:
            pred[...] = 0

        def call(dec, lo, hi):
            N.check(N.lib().qb_sw_decode(dec._h, N.ptr(det[lo:hi]), hi - lo, N.ptr(pred[lo:hi]), C.byref(dec.stats)))

        self._split(det.shape[0], call)
        return pred

    def decode_packed(self, det_rows: np.ndarray) -> np.ndarray:
        det_rows = np.ascontiguousarray(det_rows, dtype=np.uint64)
        out = np.zeros((det_rows.shape[0], max(1, (self.K + 63) // 64)), dtype=np.uint64)

        def call(dec, lo, hi):
            N.check(N.lib().qb_sw_decode_packed(dec._h, N.ptr(det_rows[lo:hi]), hi - lo, N.ptr(out[lo:hi]), C.byref(dec.stats)))

        self._split(det_rows.shape[0], call)
        return out


class MonteCarlo:
    """sample -> decode -> count on the device(s); ``run`` covers global shots [shot0, shot0 + shots).

    With an explicit context everything runs on that device (one process per GPU: bench.py, ``run_sharded``); without one the
    shots of a call are spread over every device of ``quits_b200.devices``."""

    def __init__(self, circuit, m: int, W: int, F: int, ctx: Context = None, **bp_kwargs):
        self.decoder = SlidingWindowDecoder(circuit, m, W, F, ctx=ctx, **bp_kwargs)
        self.ctx = self.decoder.ctx
        self.circuit = self.decoder.circuit
        self.K = self.decoder.K

    def run(self, shots: int, seed: int, shot0: int = 0):
        """Returns (counts uint64 [1+K], stats dict): counts[0] = shots with any observable mispredicted."""
        from . import devices as dv
        shots, shot0 = int(shots), int(shot0)

        def call(slot, lo, hi):
            dec = self.decoder._peer(slot)
            counts = np.zeros(1 + self.K, dtype=np.uint64)
            st = N.QbStats()
            N.check(N.lib().qb_mc_run(dec.ctx._h, self.circuit.for_slot(slot)._h, dec._h, int(seed), shot0 + lo, hi - lo, N.ptr(counts),
                                      C.byref(st)))
            return counts, st.as_dict()

        if not self.decoder._fanout or shot0 % 64:
            return call(0, 0, shots)
        for slot, _, _ in dv.plan_split(shots):
            self.decoder._peer(slot)
        parts = dv.run_split(shots, call)
        counts, stats = parts[0]
        for c, st in parts[1:]:
            counts = counts + c
            for k, v in st.items():                   # times: the slowest device; everything else adds up
                stats[k] = max(stats[k], v) if k.endswith("_ms") or k == "osd_max_columns" else stats[k] + v
        return counts, stats


def shard_range(total: int, rank: int, world: int, align: int = 64):
    """Global shot range of ``rank``: contiguous, aligned to 64-shot words, union = [0, total)."""
    words = (total + align - 1) // align
    lo = (words * rank // world) * align
    hi = min(total, (words * (rank + 1) // world) * align)
    return lo, max(lo, hi)


def run_sharded(circuit, m, W, F, shots, seed, rank=None, world=None, **bp_kwargs):
    """One process per GPU: this rank's share of ``shots``; results are identical for any world size because the
    noise is a function of (seed, global shot index).  Returns (counts, stats, (lo, hi)); reduce counts by summing."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    lo, hi = shard_range(int(shots), rank, world)
    mc = MonteCarlo(circuit, m, W, F, **bp_kwargs)
    counts, stats = mc.run(hi - lo, seed, lo)
    return counts, stats, (lo, hi)
