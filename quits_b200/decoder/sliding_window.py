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
from scipy.sparse import csc_matrix, hstack as sp_hstack, identity, kron as sp_kron

from ..circuit import Circuit
from ..engine import SlidingWindowDecoder
from .base import WindowPlan
from .inner import BpLsdDecoder, BpOsdDecoder, _GpuInnerDecoder, lsd_engine_options


# Decoders are set up once per (circuit text, window geometry, options) and reused by later calls of the drop-in functions:
# the reference rebuilds its ldpc decoders on every call (sliding_window.py:146-153), here that would mean re-analysing the
# circuit and re-allocating the device batch buffers each time.  Small LRU; keyed by content, not identity.
_DECODER_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_PHENOM_CACHE: "collections.OrderedDict" = collections.OrderedDict()
_DECODER_CACHE_SIZE = 4


def _freeze(v):
    """Hashable image of an option value; arrays by content (numpy's repr elides the middle of long arrays)."""
    if isinstance(v, np.ndarray):
        return ("ndarray", v.dtype.str, v.shape, hashlib.sha1(np.ascontiguousarray(v).tobytes()).hexdigest())
    if isinstance(v, (list, tuple)):
        return (type(v).__name__,) + tuple(_freeze(x) for x in v)
    return repr(v)


def _options_key(kw: dict):
    return tuple(sorted((k, _freeze(v)) for k, v in kw.items()))


def _device_key():
    from .. import devices as dv
    return tuple(dv.active_devices())


def clear_decoder_cache():
    """Drop the cached sliding-window decoders (and with them their device buffers)."""
    _DECODER_CACHE.clear()
    _PHENOM_CACHE.clear()


def _cached_decoder(circuit, m, W, F, num_cor_rounds, kw) -> SlidingWindowDecoder:
    c = Circuit.of(circuit)
    key = (_device_key(), hashlib.sha1(c.text.encode()).hexdigest(), int(m), int(W), int(F), int(num_cor_rounds), _options_key(kw))
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
        kw = lsd_engine_options(kw)
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
    gpu_classes = all(isinstance(cls, type) and issubclass(cls, _GpuInnerDecoder) for cls in (decoder1, decoder2))
    fused = gpu_classes and function_name1 == "decode" and function_name2 == "decode"
    if fused:
        kw1 = _engine_kwargs(decoder1, dict1, error_rate_name1)
        kw2 = _engine_kwargs(decoder2, dict2, error_rate_name2)
        fused = _options_key(kw1) == _options_key(kw2)
    if not fused:
        # the plug-in seam as the reference defines it (sliding_window.py:146-153,171,182): any decoder class, different classes or
        # options for the last window, any decode-function name.  Windows one after the other; the engine's own inner decoders
        # take all shots of a window in one batch, a foreign decoder is called once per shot like the reference calls it.
        return _plugin_window_loop(zcheck_samples, circuit, hz, lz, W, F, num_cor_rounds, decoder1, decoder2, dict1, dict2,
                                   error_rate_name1, error_rate_name2, function_name1, function_name2)
    dec = _cached_decoder(circuit, m, W, F, num_cor_rounds, kw1)
    # the reference leaves the priors of the last constructed decoders in the caller's dicts (sliding_window.py:148,151)
    if dec.plan.n_windows > 1:
        dict1[error_rate_name1] = dec.plan.window(dec.plan.n_windows - 2)["priors"]
    dict2[error_rate_name2] = dec.plan.window(dec.plan.n_windows - 1)["priors"]
    return dec.decode(zcheck_samples)


def _decode_window(dec, fname, syn):
    """All shots of one window through one inner decoder: ehat uint8 [N, n]."""
    if isinstance(dec, _GpuInnerDecoder) and fname == "decode":
        return dec.decode_batch(syn, want_llr=False)[0]
    fn = getattr(dec, fname)
    return np.stack([np.asarray(fn(syn[i])).astype(np.uint8) % 2 for i in range(syn.shape[0])]) if syn.shape[0] else \
        np.zeros((0, 0), dtype=np.uint8)


def _window_loop(z, m, windows, decs, fnames):
    """The reference's window loop (sliding_window.py:162-186 / :72-99) with the shot loop inside: windows = dicts with row0, H,
    L (observables x committed columns), U (carry rows x committed columns, None for the last window)."""
    n = z.shape[0]
    acc = np.zeros((n, windows[0]["L"].shape[0]), dtype=np.int64)
    carry = np.zeros((n, m), dtype=np.int64)
    for k, w in enumerate(windows):
        rows = w["H"].shape[0]
        syn = z[:, w["row0"]:w["row0"] + rows].astype(np.int64)
        syn[:, :m] = (syn[:, :m] + carry) % 2
        ehat = _decode_window(decs[k], fnames[k], syn.astype(np.uint8))
        e = ehat[:, :w["L"].shape[1]].astype(np.int64)
        acc = (acc + np.asarray((w["L"] @ e.T).T)) % 2
        if w.get("U") is not None:
            carry = np.asarray((w["U"] @ e.T).T) % 2
    return np.asarray(acc, dtype=np.int64)


def _plugin_window_loop(zcheck_samples, circuit, hz, lz, W, F, num_cor_rounds, decoder1, decoder2, dict1, dict2, error_rate_name1,
                        error_rate_name2, function_name1, function_name2):
    from .base import spacetime
    m = hz.shape[0]
    checks, observables, priors, updates = spacetime(circuit, hz, W, F, num_cor_rounds)
    decs = []
    for k in range(len(checks) - 1):
        dict1[error_rate_name1] = priors[k]
        decs.append(decoder1(checks[k], **dict1))
    dict2[error_rate_name2] = priors[-1]
    decs.append(decoder2(checks[-1], **dict2))
    windows = [{"row0": F * k * m, "H": checks[k], "L": observables[k], "U": updates[k] if k < num_cor_rounds else None}
               for k in range(num_cor_rounds + 1)]
    z = np.asarray(zcheck_samples) % 2
    return _window_loop(z, m, windows, decs, [function_name1] * num_cor_rounds + [function_name2])


def _phenom_plan(hz, lz, W, F, num_cor_rounds, W_last, D, priors1, priors2) -> WindowPlan:
    """Windows of the phenomenological model (reference sliding_window.py:56-69): H = [I_W (x) hz | B (x) I_m] with B the
    lower bidiagonal W x W matrix; the last window drops the final measurement-error block (its last round is ideal).
    Commit and carry follow :86-88 -- correction = sum of the first F data-error blocks, carry = the F-th measurement-error
    block -- written as the (L, U) pair the kernels apply: L[:, t n + q] = lz[:, q] for t < F, U[r, W n + (F-1) m + r] = 1."""
    hz = csc_matrix(np.asarray(hz) % 2, dtype=np.uint8)
    lz = csc_matrix(np.asarray(lz) % 2, dtype=np.uint8)
    m, n = hz.shape
    K = lz.shape[0]

    def window(k, Wk, last, priors):
        nm = Wk - 1 if last else Wk
        B = np.eye(Wk, dtype=np.uint8)
        for i in range(1, Wk):
            B[i, i - 1] = 1
        H = sp_hstack([sp_kron(identity(Wk, dtype=np.uint8, format="csc"), hz, format="csc"),
                       sp_kron(csc_matrix(B[:, :nm]), identity(m, dtype=np.uint8, format="csc"), format="csc")], format="csc")
        ncols = Wk * n + nm * m
        ncommit_blocks = Wk if last else F
        L = sp_hstack([lz] * ncommit_blocks + [csc_matrix((K, ncols - ncommit_blocks * n), dtype=np.uint8)], format="csc")
        U = None
        if not last:
            U = csc_matrix((np.ones(m, dtype=np.uint8), (np.arange(m), Wk * n + (F - 1) * m + np.arange(m))), shape=(m, ncols))
        pr = np.asarray(priors, dtype=np.float64)
        pr = np.full(ncols, float(pr)) if pr.ndim == 0 else pr
        if pr.shape != (ncols,):
            raise ValueError("need one prior per column of the window matrix (%d), got %r" % (ncols, pr.shape))
        return {"row0": F * k * m, "H": H, "priors": pr, "L": L, "U": U}

    windows = [window(k, W, False, priors1) for k in range(num_cor_rounds)]
    windows.append(window(num_cor_rounds, W_last, True, priors2))
    return windows if D is None else WindowPlan.explicit(m, K, D, windows)


def _phenom_priors(params: dict):
    for name in ("error_channel", "channel_probs", "error_rate"):
        if params.get(name) is not None:
            return params[name]
    raise ValueError("the inner decoder needs error_rate, error_channel or channel_probs")


def sliding_window_phenom_mem(zcheck_samples, hz, lz, W, F, decoder1, decoder2, dict1: dict, dict2: dict,
                              function_name1: str, function_name2: str, tqdm_on=False):
    """Phenomenological sliding window (reference ``sliding_window.py:14-101``) on the batched GPU kernels: same window
    matrices, commit rule and carry rule; all shots of the batch are decoded together."""
    if F == 0:
        raise ValueError("Input parameter F cannot be zero.")
    zcheck_samples = np.asarray(zcheck_samples)
    hz = np.asarray(hz)
    lz = np.asarray(lz)
    m = hz.shape[0]
    num_rounds = zcheck_samples.shape[1] // m - 2
    if 2 + num_rounds - W >= 0:
        num_cor_rounds = (2 + num_rounds - W) // F
        if (2 + num_rounds - W) % F != 0:
            num_cor_rounds += 1
    else:
        num_cor_rounds = 0
        warnings.warn("Window size larger than the syndrome extraction rounds: Doing whole history correction")
    W_last = num_rounds + 2 - F * num_cor_rounds
    if num_cor_rounds and F > W:
        raise ValueError("cannot reshape the first F data-error blocks out of a window of W < F rounds")
    rate_names = ("error_rate", "error_channel", "channel_probs")
    gpu_classes = all(isinstance(cls, type) and issubclass(cls, _GpuInnerDecoder) for cls in (decoder1, decoder2))
    fused = gpu_classes and function_name1 == "decode" and function_name2 == "decode"
    if fused:
        kw1 = _engine_kwargs(decoder1, {k: v for k, v in dict1.items() if k not in rate_names}, "")
        kw2 = _engine_kwargs(decoder2, {k: v for k, v in dict2.items() if k not in rate_names}, "")
        fused = _options_key(kw1) == _options_key(kw2)
    if not fused:
        # the plug-in seam (sliding_window.py:56-69,84,94): any decoder classes / options / function names; the window matrices are
        # the reference's, the inner decoders are constructed and called as the reference constructs and calls them
        wins = _phenom_plan(hz, lz, W, F, num_cor_rounds, W_last, None, 0.5, 0.5)
        decs = [decoder1(w["H"], **dict1) for w in wins[:-1]] + [decoder2(wins[-1]["H"], **dict2)]
        z = zcheck_samples[:, :m * (num_rounds + 2)] % 2
        return _window_loop(z, m, wins, decs, [function_name1] * num_cor_rounds + [function_name2])
    p1, p2 = _phenom_priors(dict1), _phenom_priors(dict2)
    key = (_device_key(), hashlib.sha1(np.ascontiguousarray(hz % 2, dtype=np.uint8).tobytes()).hexdigest(), hz.shape,
           hashlib.sha1(np.ascontiguousarray(lz % 2, dtype=np.uint8).tobytes()).hexdigest(), int(W), int(F), int(num_rounds),
           hashlib.sha1(np.asarray(p1, dtype=np.float64).tobytes()).hexdigest(), hashlib.sha1(np.asarray(p2, dtype=np.float64).tobytes()).hexdigest(),
           _options_key(kw1))
    dec = _PHENOM_CACHE.get(key)
    if dec is None:
        plan = _phenom_plan(hz, lz, W, F, num_cor_rounds, W_last, m * (num_rounds + 2), p1, p2)
        dec = SlidingWindowDecoder.from_plan(plan, **kw1)
        _PHENOM_CACHE[key] = dec
        while len(_PHENOM_CACHE) > _DECODER_CACHE_SIZE:
            _PHENOM_CACHE.popitem(last=False)
    else:
        _PHENOM_CACHE.move_to_end(key)
    return dec.decode(zcheck_samples[:, :m * (num_rounds + 2)])


__all__ = ["sliding_window_phenom_mem", "sliding_window_circuit_mem", "clear_decoder_cache"]
