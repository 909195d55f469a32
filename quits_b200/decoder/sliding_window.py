"""Sliding-window decoding over the circuit-level detector error matrix, batched on the GPU.

Same call signature, return dtype, warning and exceptions as the reference's
``sliding_window_circuit_mem`` (reference ``src/quits/decoder/sliding_window.py:104-188``); the per-shot Python
loop (:162-186) is replaced by the CUDA kernels in quits_b200/csrc (one launch sequence per batch of shots).
"""
from __future__ import annotations

import collections
import hashlib
import warnings

import numpy as np

from ..circuit import Circuit
from ..engine import SlidingWindowDecoder
from .inner import BpLsdDecoder, BpOsdDecoder, _GpuInnerDecoder


# Decoders are set up once per (circuit text, window geometry, options) and reused by later calls of the drop-in functions:
# the reference rebuilds its ldpc decoders on every call (sliding_window.py:146-153), here that would mean re-analysing the
# circuit and re-allocating the device batch buffers each time.  Small LRU; keyed by content, not identity.
_DECODER_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_DECODER_CACHE_SIZE = 4


def _cached_decoder(circuit, m, W, F, num_cor_rounds, kw) -> SlidingWindowDecoder:
    c = Circuit.of(circuit)
    key = (hashlib.sha1(c.text.encode()).hexdigest(), int(m), int(W), int(F), int(num_cor_rounds),
           tuple(sorted((k, repr(v)) for k, v in kw.items())))
    dec = _DECODER_CACHE.get(key)
    if dec is None:
        dec = SlidingWindowDecoder(c, m, W, F, num_cor_rounds, **kw)
        _DECODER_CACHE[key] = dec
        while len(_DECODER_CACHE) > _DECODER_CACHE_SIZE:
            _DECODER_CACHE.popitem(last=False)
    else:
        _DECODER_CACHE.move_to_end(key)
    return dec


def _engine_kwargs(decoder_cls, params: dict, rate_name: str) -> dict:
    kw = {k: v for k, v in params.items() if k != rate_name}
    if issubclass(decoder_cls, BpLsdDecoder):
        method = str(kw.pop("lsd_method", "lsd_0")).lower()
        kw.pop("lsd_order", None)
        if method not in ("off", "none"):
            raise NotImplementedError("BP-LSD post-processing (lsd_method=%r) is not implemented on the GPU path yet" % method)
        kw["osd_method"] = "off"
    return kw


def sliding_window_circuit_mem(zcheck_samples, circuit, hz, lz, W, F, decoder1, decoder2, dict1: dict, dict2: dict,
                               error_rate_name1: str, error_rate_name2: str, function_name1: str, function_name2: str,
                               tqdm_on=False):
    """Sliding-window decoder (Huang & Puri, PRA 110, 012453) on the spacetime detector error matrix.

    ``decoder1`` / ``decoder2`` must be the engine's GPU inner decoders (``quits_b200.decoder.BpOsdDecoder`` /
    ``BpLsdDecoder``, or the classes ``quits_b200.compat`` registers under ldpc's names); all shots are then decoded
    by the batched kernels.  Returns ``int64 [num_trials, K]`` like the reference (sliding_window.py:160).
    """
    zcheck_samples = np.asarray(zcheck_samples)
    m = hz.shape[0]
    num_rounds = zcheck_samples.shape[1] // m - 2
    if 2 + num_rounds - W >= 0:
        num_cor_rounds = (2 + num_rounds - W) // F
        if (2 + num_rounds - W) % F != 0:
            num_cor_rounds += 1
    else:
        num_cor_rounds = 0
        warnings.warn("Window size larger than the syndrome extraction rounds: Doing whole history correction")
    if F == 0:
        raise ValueError("Input parameter F cannot be zero.")
    for cls in (decoder1, decoder2):
        if not (isinstance(cls, type) and issubclass(cls, _GpuInnerDecoder)):
            raise NotImplementedError(
                "the GPU sliding-window path runs the engine's own inner decoders only (quits_b200.decoder.BpOsdDecoder / "
                "BpLsdDecoder); got %r. There is no per-shot CPU fallback." % (cls,))
    kw1 = _engine_kwargs(decoder1, dict1, error_rate_name1)
    kw2 = _engine_kwargs(decoder2, dict2, error_rate_name2)
    if kw1 != kw2 or function_name1 != "decode" or function_name2 != "decode":
        raise NotImplementedError("different inner decoders for the sliding windows and the last window are not supported on the GPU path")
    dec = _cached_decoder(circuit, m, W, F, num_cor_rounds, kw1)
    # the reference leaves the priors of the last constructed decoders in the caller's dicts (sliding_window.py:148,151)
    if dec.plan.n_windows > 1:
        dict1[error_rate_name1] = dec.plan.window(dec.plan.n_windows - 2)["priors"]
    dict2[error_rate_name2] = dec.plan.window(dec.plan.n_windows - 1)["priors"]
    return dec.decode(zcheck_samples)


def sliding_window_phenom_mem(zcheck_samples, hz, lz, W, F, decoder1, decoder2, dict1: dict, dict2: dict,
                              function_name1: str, function_name2: str, tqdm_on=False):
    """Phenomenological sliding window (reference sliding_window.py:14-101): not on the GPU path yet."""
    raise NotImplementedError("sliding_window_phenom_mem is not implemented on the GPU path yet (circuit-level windows only)")


__all__ = ["sliding_window_phenom_mem", "sliding_window_circuit_mem"]
