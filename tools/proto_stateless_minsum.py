#!/usr/bin/env python
"""Prototype (CPU, pure Python) of the message-free form of flooding min-sum that DESIGN.md section 7 proposes for windows whose
messages do not fit in shared memory (BASELINE config 5: 150 830 edges = 1.2 MB of fp64 messages per shot).

Flooding min-sum as ldpc runs it (oracle/bp_impl.inc:40-78) keeps one bit->check message v_e per edge.  A check->bit message only
needs, of its row, the two smallest magnitudes (m1 <= m2), the row's sign parity, and of the edge itself TWO BITS: s_e = [v_e <= 0]
and f_e = [|v_e| == m1]:

    c_e = (f_e ? m2 : m1) * ((parity + s_e) even ? alpha : -alpha)            (ties: |v_e| == m1 twice  =>  m2 == m1)

and the next messages v_e = (l0 + c_0 + .. + c_{q-1}) + (0 + c_{W-1} + .. + c_{q+1}) are consumed immediately: by the posterior, and by
the NEXT row summaries (min1 / min2 / parity of the new magnitudes and signs, an atomic-min style reduction per row) and the new
two bits per edge.  State per shot: 2 bits per edge + (m1, m2, parity) per row instead of 64 bits per edge -- 37 KB + 36 KB for a
config-5 window, shared-memory resident.  The price is that f_e needs the FINAL m1 of the row, so the bit sweep runs twice per
iteration (magnitudes first, flags second) -- arithmetic instead of 3 x 1.2 MB of scattered HBM traffic per iteration.

This file proves the arithmetic: `decode` returns the same posteriors, bit for bit, as the oracle's flooding min-sum.

    python tools/proto_stateless_minsum.py          # self-check on random matrices
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIG = sys.float_info.max


def decode(H, priors, syn, max_iter, alpha_opt=1.0):
    """H dense uint8 [m][n].  Returns (ehat, llr, iterations, converged) of flooding min-sum, never holding a message array."""
    m, n = H.shape
    cols = [np.flatnonzero(H[:, j]) for j in range(n)]                 # rows of column j, ascending (the oracle's edge order)
    llr0 = [math.log((1.0 - p) / p) for p in priors]
    # state: per row (m1, m2, parity incl. syndrome); per edge (column j, position q): s, f
    m1 = [BIG] * m; m2 = [BIG] * m; par = [int(syn[i]) & 1 for i in range(m)]
    s = [[0] * len(cols[j]) for j in range(n)]
    f = [[0] * len(cols[j]) for j in range(n)]

    def fold(rows_m1, rows_m2, i, a):                                   # two smallest of a row, one element at a time
        if a < rows_m1[i]:
            rows_m2[i] = rows_m1[i]; rows_m1[i] = a
        elif a < rows_m2[i]:
            rows_m2[i] = a
    for j in range(n):                                                  # iteration 0 messages are the priors
        for q, i in enumerate(cols[j]):
            fold(m1, m2, i, abs(llr0[j]))
            s[j][q] = 1 if llr0[j] <= 0 else 0
            par[i] += s[j][q]
    for j in range(n):
        for q, i in enumerate(cols[j]):
            f[j][q] = 1 if abs(llr0[j]) == m1[i] else 0
    llr = list(llr0)
    ehat = np.zeros(n, np.uint8)
    for it in range(1, max_iter + 1):
        alpha = (1.0 - 2.0 ** (-it)) if alpha_opt == 0.0 else alpha_opt

        def new_messages(j):
            c = []
            for q, i in enumerate(cols[j]):
                mag = m2[i] if f[j][q] else m1[i]
                c.append(mag * (alpha if (par[i] + s[j][q]) % 2 == 0 else -alpha))
            W = len(c)
            pre, t = [], llr0[j]
            for q in range(W):
                pre.append(t); t = t + c[q]
            post = t
            v, t = [0.0] * W, 0.0
            for q in range(W - 1, -1, -1):
                v[q] = pre[q] + t; t = t + c[q]
            return v, post
        n1 = [BIG] * m; n2 = [BIG] * m; npar = [int(syn[i]) & 1 for i in range(m)]
        ns = [[0] * len(cols[j]) for j in range(n)]
        cand = [0] * m
        for j in range(n):                                              # sweep 1: posteriors, new magnitudes and signs into the rows
            v, post = new_messages(j)
            llr[j] = post
            ehat[j] = 1 if post <= 0 else 0
            for q, i in enumerate(cols[j]):
                if ehat[j]:
                    cand[i] ^= 1
                fold(n1, n2, i, abs(v[q]))
                ns[j][q] = 1 if v[q] <= 0 else 0
                npar[i] += ns[j][q]
        if all(cand[i] == (int(syn[i]) & 1) for i in range(m)):
            return ehat, np.array(llr), it, True
        nf = [[0] * len(cols[j]) for j in range(n)]
        for j in range(n):                                              # sweep 2: the same messages again, now against the final minima
            v, _ = new_messages(j)
            for q, i in enumerate(cols[j]):
                nf[j][q] = 1 if abs(v[q]) == n1[i] else 0
        m1, m2, par, s, f = n1, n2, npar, ns, nf
    return ehat, np.array(llr), max_iter, False


def self_check(trials=60, seed=1):
    import scipy.sparse as sp
    from oracle import cref
    rng = np.random.default_rng(seed)
    checked = 0
    for t in range(trials):
        m = int(rng.integers(4, 24)); n = int(rng.integers(m, 3 * m + 4))
        H = (rng.random((m, n)) < min(0.5, 3.0 / m)).astype(np.uint8)
        for j in range(n):
            if not H[:, j].any():
                H[rng.integers(m), j] = 1
        p = rng.choice([0.01, 0.03, 0.08], n) if t % 2 else rng.uniform(0.005, 0.2, n)
        alpha = [1.0, 0.75, 0.0][t % 3]
        it = int(rng.integers(1, 9))
        orc = cref.BpOsd(sp.csc_matrix(H), p, max_iter=it, bp_method="minimum_sum", schedule="parallel", ms_scaling_factor=alpha, osd=False)
        for k in range(3):
            syn = (H @ (rng.random(n) < 0.12).astype(np.uint8) % 2).astype(np.uint8)
            e, l, iters, conv = orc.decode(syn)
            e2, l2, iters2, conv2 = decode(H, p, syn, it, alpha)
            assert iters == iters2 and bool(conv) == conv2, (t, k)
            assert np.array_equal(l, l2), (t, k, float(np.max(np.abs(l - l2))))
            assert np.array_equal(e, e2)
            checked += 1
    return checked


if __name__ == "__main__":
    print("message-free flooding min-sum equals the oracle bit for bit on %d decodes" % self_check())
