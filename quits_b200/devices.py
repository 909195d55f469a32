"""Which GPUs the drop-in calls spread over, and the plumbing that spreads them.

The reference's loop is serial over shots (``src/quits/decoder/sliding_window.py:162-186`` re-initialises per shot: shots are
i.i.d.), so a call with N shots is cut into contiguous shot ranges, one per device, each handled by a host thread that drives its
own context (ctypes releases the GIL for the duration of a C call; every C entry point selects its context's device).  Shot s of a
seed is the same bits on any device and for any split (counter-based noise, ``qb_sample``), so the result of a call does not depend
on how many devices served it.

Default: every visible device -- unless the process is one rank of a multi-process job (``WORLD_SIZE`` > 1: torchrun gives each rank
one GPU, ``LOCAL_RANK``).  ``QB_DEVICES`` ("all", "1", "0,2,3") or :func:`set_devices` override.  A device may be listed more than
once (each listing gets its own context and stream), which is how the single-GPU tests exercise the fan-out.
"""
from __future__ import annotations

import os
import threading

from . import _native as N

MIN_SHOTS_PER_DEVICE = 16384          # below this a second device costs more in set-up and launch latency than it saves

_lock = threading.Lock()
_devices = None                       # list of device indices, one entry per slot
_slot_ctx = {}                        # slot -> Context


def _default_devices():
    n = max(N.device_count(), 0)
    if n == 0:
        return [0]
    local = int(os.environ.get("LOCAL_RANK", "0")) % n
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return [local]
    spec = os.environ.get("QB_DEVICES", "all").strip().lower()
    if spec in ("", "all"):
        return [local] + [d for d in range(n) if d != local]
    devs = [int(x) % n for x in spec.split(",") if x.strip() != ""]
    return devs or [local]


def active_devices():
    """Device index of every slot the drop-in calls may use (slot 0 is the default context's device)."""
    global _devices
    with _lock:
        if _devices is None:
            _devices = _default_devices()
        return list(_devices)


def set_devices(devs):
    """Use these devices (indices; repeats allowed) for the calls that follow; ``None`` restores the default."""
    global _devices
    with _lock:
        _devices = None if devs is None else [int(d) for d in devs]
        _slot_ctx.clear()


def slot_context(slot: int):
    """The context of a slot: slot 0 shares ``Context.default()`` of its device, later slots own theirs."""
    from .circuit import Context
    devs = active_devices()
    with _lock:
        ctx = _slot_ctx.get(slot)
        if ctx is None:
            ctx = Context.default(devs[slot]) if slot == 0 else Context(devs[slot])
            _slot_ctx[slot] = ctx
        return ctx


def plan_split(n: int, align: int = 64):
    """[(slot, lo, hi)]: contiguous shot ranges aligned to 64-shot words, at least MIN_SHOTS_PER_DEVICE shots per slot used."""
    n = int(n)
    slots = max(1, min(len(active_devices()), n // MIN_SHOTS_PER_DEVICE))
    words = (n + align - 1) // align
    out = []
    for s in range(slots):
        lo = min(n, (words * s // slots) * align)
        hi = min(n, (words * (s + 1) // slots) * align)
        if hi > lo or (s == 0 and n == 0):
            out.append((s, lo, hi))
    return out or [(0, 0, n)]


def run_split(n: int, fn):
    """Call ``fn(slot, lo, hi)`` for every range of :func:`plan_split`, one host thread per slot; re-raises the first failure."""
    parts = plan_split(n)
    if len(parts) == 1:
        s, lo, hi = parts[0]
        return [fn(s, lo, hi)]
    results, errors = [None] * len(parts), []

    def work(i, s, lo, hi):
        try:
            results[i] = fn(s, lo, hi)
        except BaseException as e:      # noqa: BLE001 - re-raised below on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(i, s, lo, hi), daemon=True) for i, (s, lo, hi) in enumerate(parts)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results
