"""Oracle: circuit -> detector error model (backward sensitivity sweep), plain Python.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates what the reference obtains from
``circuit.detector_error_model(decompose_errors=False)`` (reference ``decoder/base.py:151``), i.e. the
published behaviour of Stim's ErrorAnalyzer (stim>=1.13, not vendored -- parity unpinned):

* walk the flattened circuit backwards keeping, per qubit, the set of detectors / observables flipped by an
  X resp. Z error at that point (``H`` swaps them; ``CX c t``: ``xs_c ^= xs_t``, ``zs_t ^= zs_c``; ``M``: ``xs ^= D[m]``;
  ``MX``: ``zs ^= D[m]``; ``MR``: ``xs = D[m], zs = 0``; ``R``/``RX``: clear);
* every noise channel is split into independent components: ``X_ERROR/Z_ERROR(p)`` -> p;
  ``DEPOLARIZE1(p)`` -> X, Y, Z each with ``q = 1/2 - 1/2 sqrt(1 - 4p/3)``; ``DEPOLARIZE2(p)`` -> 15 Pauli pairs
  each with ``q = 1/2 - 1/2 (1 - 16p/15)^(1/8)``;
* components with the same symptom set XOR-combine ``p <- p(1-q) + q(1-p)``; empty symptoms are dropped;
* errors are emitted in ascending lexicographic order of (sorted detector ids, then observable ids) -- the
  order the reference's ``spacetime`` slicing relies on (``decoder/base.py:163,169,178``).

Combination order (it fixes the last bits of the priors): instructions last-to-first, targets
last-to-first, Pauli codes c = 1..3 (1..15) ascending with ``c&1 = X_a, c>>1&1 = Z_a, c>>2&1 = X_b, c>>3&1 = Z_b``.
The product's host analyser (quits_b200/csrc/dem.cpp) uses the same order so priors agree bit for bit.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

from .stimtext import FlatCircuit


@dataclass
class Dem:
    n_det: int
    n_obs: int
    probs: List[float]
    dets: List[List[int]]
    obs: List[List[int]]
    rep: List[Tuple[int, int, int]]   # representative fault per error: (flat op index, target/pair index, pauli code)


def dep1_component(p: float) -> float:
    return 0.5 - 0.5 * math.sqrt(1.0 - 4.0 * p / 3.0)


def dep2_component(p: float) -> float:
    return 0.5 - 0.5 * math.pow(1.0 - 16.0 * p / 15.0, 0.125)


def _bits(x: int) -> List[int]:
    out = []
    while x:
        low = x & -x
        out.append(low.bit_length() - 1)
        x ^= low
    return out


def analyze(fc: FlatCircuit) -> Dem:
    D, K = fc.n_det, fc.n_obs
    sens = [0] * fc.n_meas                     # per measurement: bitmask of detectors (bits < D) / observables (bits D+i)
    for op in fc.ops:
        if op.name == "DETECTOR":
            bit = 1 << int(op.arg)
        elif op.name == "OBSERVABLE_INCLUDE":
            bit = 1 << (D + int(op.arg))
        else:
            continue
        for m in op.targets:
            sens[m] ^= bit
    # measurement index of the first target of every measuring op
    mbase = {}
    cnt = 0
    for i, op in enumerate(fc.ops):
        if op.name in ("M", "MX", "MR"):
            mbase[i] = cnt
            cnt += len(op.targets)

    xs = [0] * fc.n_qubits
    zs = [0] * fc.n_qubits
    prob = {}
    rep = {}

    def add(sym: int, q: float, where) -> None:
        if sym == 0 or q == 0.0:
            return
        if sym in prob:
            p0 = prob[sym]
            prob[sym] = p0 * (1.0 - q) + q * (1.0 - p0)
        else:
            prob[sym] = q
            rep[sym] = where

    for i in range(len(fc.ops) - 1, -1, -1):
        op = fc.ops[i]
        name, t = op.name, op.targets
        if name in ("DETECTOR", "OBSERVABLE_INCLUDE"):
            continue
        if name == "CX":
            for j in range(len(t) - 2, -1, -2):
                c, tg = t[j], t[j + 1]
                xs[c] ^= xs[tg]
                zs[tg] ^= zs[c]
        elif name == "H":
            for q in t:
                xs[q], zs[q] = zs[q], xs[q]
        elif name == "R":
            for q in t:
                if zs[q]:
                    raise ValueError("non-deterministic detector/observable: sensitive to Z on qubit %d right after R" % q)
                xs[q] = zs[q] = 0
        elif name == "RX":
            for q in t:
                if xs[q]:
                    raise ValueError("non-deterministic detector/observable: sensitive to X on qubit %d right after RX" % q)
                xs[q] = zs[q] = 0
        elif name == "M":
            for j in range(len(t) - 1, -1, -1):
                xs[t[j]] ^= sens[mbase[i] + j]
        elif name == "MX":
            for j in range(len(t) - 1, -1, -1):
                zs[t[j]] ^= sens[mbase[i] + j]
        elif name == "MR":
            for j in range(len(t) - 1, -1, -1):
                q = t[j]
                if zs[q]:
                    raise ValueError("non-deterministic detector/observable: sensitive to Z on qubit %d right after MR" % q)
                xs[q] = sens[mbase[i] + j]
                zs[q] = 0
        elif name == "X_ERROR":
            for j in range(len(t) - 1, -1, -1):
                add(xs[t[j]], op.arg, (i, j, 1))
        elif name == "Z_ERROR":
            for j in range(len(t) - 1, -1, -1):
                add(zs[t[j]], op.arg, (i, j, 2))
        elif name == "DEPOLARIZE1":
            q1 = dep1_component(op.arg)
            for j in range(len(t) - 1, -1, -1):
                a = t[j]
                for c in (1, 2, 3):
                    add((xs[a] if c & 1 else 0) ^ (zs[a] if c & 2 else 0), q1, (i, j, c))
        elif name == "DEPOLARIZE2":
            q2 = dep2_component(op.arg)
            for j in range(len(t) // 2 - 1, -1, -1):
                a, b = t[2 * j], t[2 * j + 1]
                for c in range(1, 16):
                    sym = ((xs[a] if c & 1 else 0) ^ (zs[a] if c & 2 else 0) ^
                           (xs[b] if c & 4 else 0) ^ (zs[b] if c & 8 else 0))
                    add(sym, q2, (i, j, c))
        else:
            raise NotImplementedError(name)

    dmask = (1 << D) - 1
    items = []
    for sym, p in prob.items():
        dets = _bits(sym & dmask)
        obs = _bits(sym >> D)
        key = tuple(dets) + tuple(D + 10**9 + o for o in obs)     # observables order after every detector
        items.append((key, p, dets, obs, rep[sym]))
    items.sort(key=lambda x: x[0])
    return Dem(D, K, [x[1] for x in items], [x[2] for x in items], [x[3] for x in items], [x[4] for x in items])
