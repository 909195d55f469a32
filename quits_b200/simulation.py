"""Drop-in for ``quits.simulation.get_stim_mem_result`` (reference ``src/quits/simulation.py:8-28``)."""
from __future__ import annotations

import numpy as np

from .circuit import Circuit


def get_stim_mem_result(circuit, num_trials, seed=-1):
    """Pauli-frame Monte-Carlo of the memory circuit on the GPU (K1 kernel).

    :param circuit: Stim circuit -- a ``quits_b200.Circuit``, Stim text, or any object whose ``str()`` is Stim text
                    (a ``stim.Circuit`` is)
    :param num_trials: number of shots
    :param seed: run seed; negative => drawn from OS entropy, as the reference does when no seed is given
    :return: (detection_events bool[num_trials, #detectors], observable_flips bool[num_trials, #observables])
    """
    c = Circuit.of(circuit)
    if seed is None or seed < 0:
        seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).view(np.uint64)[0]) & (2**63 - 1)
    return c.sample(int(num_trials), int(seed))


def get_codecap_pL(code, p, num_trials, decoder, dict, basis='Z', seed=-1, tqdm_on=False):
    """Code-capacity logical error rate (reference ``src/quits/simulation.py:31-61``), all trials decoded in one GPU batch.

    Same signature and the same use of numpy's global RNG as the reference: trial i draws ``np.random.binomial(1, p, n)``
    (one vectorised call here; numpy fills it element by element from the same stream, so the noise is identical for a given
    ``seed``).  ``decoder`` must be the engine's ``BpOsdDecoder`` / ``BpLsdDecoder`` class; ``dict`` holds its keyword arguments.
    """
    from .decoder.inner import _GpuInnerDecoder
    if seed >= 0:
        np.random.seed(seed)
    basis = basis.upper()
    if basis == 'Z':
        parity_check_matrix, logical_codewords = code.hz, code.lz
    elif basis == 'X':
        parity_check_matrix, logical_codewords = code.hx, code.lx
    else:
        raise ValueError("basis must be 'Z' or 'X'")
    if not (isinstance(decoder, type) and issubclass(decoder, _GpuInnerDecoder)):
        raise NotImplementedError("get_codecap_pL on the GPU path needs quits_b200.decoder.BpOsdDecoder / BpLsdDecoder, got %r; "
                                  "there is no per-shot CPU fallback" % (decoder,))
    H = np.asarray(parity_check_matrix) % 2
    Lg = np.asarray(logical_codewords) % 2
    bpd = decoder(H, **dict)
    num_trials = int(num_trials)
    num_errors = 0
    Ht, Lt = H.T.astype(np.int64), Lg.T.astype(np.int64)
    for lo in range(0, num_trials, 65536):                   # chunks: memory stays O(chunk x n) whatever num_trials is
        n = min(65536, num_trials - lo)
        noise = np.random.binomial(1, p, (n, H.shape[1])).astype(np.uint8)      # same stream as one draw of all trials
        syndromes = (noise.astype(np.int64) @ Ht % 2).astype(np.uint8)
        decoded, _, _, _ = bpd.decode_batch(syndromes, want_llr=False)
        residual = decoded ^ noise
        num_errors += int(np.any(residual.astype(np.int64) @ Lt % 2, axis=1).sum())
    return num_errors / num_trials


def count_logical_errors(observable_flips, logical_pred, threads=None):
    """Number of shots whose prediction differs from the observed flips in any observable -- the caller's reduction
    ``np.sum(np.any((observable_flips - logical_pred) % 2, axis=1))`` (reference ``tests/test_sliding_window.py:83``,
    ``doc/06B`` cell 5) without the int64 temporaries, over row blocks on a thread pool (numpy releases the GIL in these loops):
    at eight GPUs the single-threaded idiom costs more than the decode it follows."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    obs = np.asarray(observable_flips)
    pred = np.asarray(logical_pred)
    n = obs.shape[0]
    if pred.shape != obs.shape:
        raise ValueError("observable flips %r and predictions %r differ in shape" % (obs.shape, pred.shape))
    threads = max(1, min(int(threads or os.cpu_count() or 1), 32, (n + 65535) // 65536))

    def bits(a):                                      # 0/1 bytes of a boolean / integer array (integers reduced mod 2)
        if a.dtype == np.bool_:
            return a.view(np.uint8)
        if a.dtype.kind in "iu":
            return a.astype(np.uint8) & np.uint8(1)     # truncation keeps the low byte, whose low bit is the parity
        return (a % 2 != 0).view(np.uint8)

    def block(i):
        lo, hi = n * i // threads, n * (i + 1) // threads
        x = np.ascontiguousarray(bits(obs[lo:hi]) ^ bits(pred[lo:hi]))          # [rows, K] bytes, non-zero where they differ
        k = x.shape[1]
        if k == 0 or hi == lo:
            return 0
        if k % 4 == 0:                                # OR the row's bytes through a few wide column passes (an axis-1 reduction
            x = x.view(np.uint32)                     # over K = 12 short elements is an order of magnitude slower)
        acc = x[:, 0].copy()
        for c in range(1, x.shape[1]):
            acc |= x[:, c]
        return int(np.count_nonzero(acc))

    if threads == 1:
        return block(0)
    with ThreadPoolExecutor(threads) as ex:
        return int(sum(ex.map(block, range(threads))))


__all__ = ["get_stim_mem_result", "get_codecap_pL", "count_logical_errors"]
