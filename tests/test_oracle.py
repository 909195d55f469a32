"""CPU tests of the ORACLE (test infrastructure): pinned against the only bit-exact artefact the reference holds for
this path (the printed [[72,12,6]] circuit of doc/02A), against Random123's Philox known answers, and cross-checked
internally (forward frame propagation <-> backward DEM analysis; BP/OSD output properties; the reference's own
window loop, whose outputs are the committed fixtures, against the oracle's C restatement of that loop)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import BP_KW, DECODE_CASES, GOLDEN, case_circuit, circuit_text, decode_case
from oracle import cref, dem as odem, stimtext


def test_doc02A_printout_is_reproduced():
    """reference doc/02A_custom_circuit_generation.ipynb:82-289 prints stim's canonical form of the [[72,12,6]] custom
    circuit (15 rounds, p=1e-3); the oracle's parser + canonical printer must give the same 206 lines from the text the
    reference's own BbCode.build_circuit emitted (tests/golden/circuits/bb72_r15_p1e-3.stim)."""
    with open(os.path.join(GOLDEN, "ref_doc02A_bb72_r15_printout.txt")) as f:
        want = [ln.rstrip() for ln in f.read().strip("\n").split("\n")]
    got = stimtext.canonical_text(circuit_text("bb72_r15_p1e-3")).split("\n")
    assert len(want) == 206
    assert got == want


def test_canonical_text_is_the_same_circuit():
    text = circuit_text("bb72_r6_p1e-3")
    a = stimtext.parse_flat(text)
    b = stimtext.parse_flat(stimtext.canonical_text(text))
    assert (a.n_qubits, a.n_meas, a.n_det, a.n_obs) == (b.n_qubits, b.n_meas, b.n_det, b.n_obs)
    da, db = odem.analyze(a), odem.analyze(b)
    assert da.dets == db.dets and da.obs == db.obs and da.probs == db.probs


def test_survey_sizes():
    """SURVEY.md appendix C (probe of the reference): gross code, 10 rounds."""
    fc = stimtext.parse_flat(circuit_text("bb144_r10_p1e-3"))
    assert (fc.n_qubits, fc.n_meas, fc.n_det, fc.n_obs) == (288, 1728, 864, 12)
    d = odem.analyze(fc)
    assert len(d.probs) == 8064 and sum(len(x) for x in d.dets) == 28152
    assert abs(sum(d.probs) - 14.67) < 0.01


@pytest.mark.parametrize("key,ctr,want", [
    ((0, 0), (0, 0, 0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff), (0xffffffff,) * 4, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0xa4093822, 0x299f31d0), (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
])
def test_philox_known_answers(key, ctr, want):
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    assert tuple(int(x) for x in cref.philox(key[0], key[1], ctr)) == want


def test_noise_tables():
    for p in (1e-4, 1e-3, 1e-2, 0.2):
        t1, c = cref.noise_tables(p)
        assert abs(t1 / 2**32 - (1 - (1 - p) ** 64)) < 1e-9
        assert all(int(c[k]) <= int(c[k + 1]) for k in range(1, 62))
        # P(N = 1 | N >= 1)
        want = 64 * p * (1 - p) ** 63 / (1 - (1 - p) ** 64)
        assert abs(int(c[1]) / 2**64 - want) < 1e-12
    assert cref.noise_tables(0.0)[0] == 0


@pytest.mark.parametrize("name", ["bb72_r6_p1e-3", "bb72_r3_p1e-3_X", "toric3_zxcol_r3_p1e-3", "hgp225_r3_p1e-2"])
def test_forward_frame_reproduces_every_dem_column(name):
    """Inject each error's representative fault into the frame simulator: the detectors / observables that fire must be
    exactly that error's symptom (frame rules and backward analyser are two independent restatements)."""
    fc = stimtext.parse_flat(circuit_text(name))
    d = odem.analyze(fc)
    ops = np.array([r[0] for r in d.rep]); tg = np.array([r[1] for r in d.rep]); cd = np.array([r[2] for r in d.rep])
    det, obs = cref.inject(fc, ops, tg, cd)
    for i in range(len(d.probs)):
        assert np.flatnonzero(det[i]).tolist() == d.dets[i], i
        assert np.flatnonzero(obs[i]).tolist() == d.obs[i], i


def test_sampler_marginals_match_dem():
    fc = stimtext.parse_flat(circuit_text("bb72_r6_p3e-3"))
    d = odem.analyze(fc)
    shots = 64 * 400
    det, obs = cref.sample(fc, 99, 0, shots)
    prod = np.ones(fc.n_det)
    for p, ds in zip(d.probs, d.dets):
        for x in ds:
            prod[x] *= 1 - 2 * p
    want = 0.5 * (1 - prod)
    got = det.mean(axis=0)
    sigma = np.sqrt(want * (1 - want) / shots)
    assert np.all(np.abs(got - want) < 5.5 * sigma + 1e-9)
    # no-noise circuits fire nothing
    quiet = stimtext.parse_flat(circuit_text("bb72_r6_p3e-3").replace("(0.0030000000)", "(0)"))
    det0, obs0 = cref.sample(quiet, 1, 0, 128)
    assert det0.sum() == 0 and obs0.sum() == 0


@pytest.mark.parametrize("case", sorted(f[:-5] for f in os.listdir(os.path.join(GOLDEN, "windows"))))
def test_oracle_windows_equal_reference_spacetime(case):
    """oracle/windows.py against the digests of the reference's own spacetime() output (tools/make_golden.py)."""
    import hashlib
    import json
    from oracle import windows as owin
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    with open(os.path.join(GOLDEN, "windows", case + ".json")) as f:
        g = json.load(f)
    d = odem.analyze(stimtext.parse_flat(circuit_text(g["circuit"])))
    wins = owin.plan(d, g["m"], g["W"], g["F"])
    assert len(wins) == len(g["windows"])
    for w, e in zip(wins, g["windows"]):
        H = w["H"]; H.sort_indices()
        L = w["L"]; L.sort_indices()
        assert list(H.shape) == e["H_shape"] and sha(H.indptr.astype(np.int64)) == e["H_indptr"] and sha(H.indices.astype(np.int32)) == e["H_indices"]
        assert list(L.shape) == e["L_shape"] and sha(L.indptr.astype(np.int64)) == e["L_indptr"] and sha(L.indices.astype(np.int32)) == e["L_indices"]
        assert sha(np.asarray(w["priors"], dtype=np.float64)) == e["priors"]
        if "U_shape" in e:
            U = w["U"]; U.sort_indices()
            assert list(U.shape) == e["U_shape"] and sha(U.indptr.astype(np.int64)) == e["U_indptr"] and sha(U.indices.astype(np.int32)) == e["U_indices"]


def test_bp_osd_output_properties():
    fc = stimtext.parse_flat(circuit_text("bb72_r6_p3e-3"))
    d = odem.analyze(fc)
    D, C = fc.n_det, len(d.probs)
    rows = np.array([x for ds in d.dets for x in ds]); cols = np.array([j for j, ds in enumerate(d.dets) for _ in ds])
    H = sp.csc_matrix((np.ones(len(rows), dtype=np.uint8), (rows, cols)), shape=(D, C))
    pri = np.array(d.probs)
    rng = np.random.default_rng(3)
    dec64 = cref.BpOsd(H, pri, max_iter=10, bp_method="minimum_sum", schedule="parallel", precision="f64")
    dec32 = cref.BpOsd(H, pri, max_iter=10, bp_method="minimum_sum", schedule="parallel", precision="f32")
    n_osd = 0
    for _ in range(40):
        e = (rng.random(C) < pri * 1.5).astype(np.uint8)
        s = (H @ e) % 2
        e64, l64, it64, c64 = dec64.decode(s)
        e32, l32, it32, c32 = dec32.decode(s)
        assert np.array_equal((H @ e64) % 2, s) and np.array_equal((H @ e32) % 2, s)      # BP-converged or OSD: always a solution
        # fp32 and fp64 follow the same trajectory unless an exactly-tied LLR (common with 9 distinct priors) resolves
        # differently; when they do, the posteriors agree to fp32 rounding accumulated over the iterations
        if it64 == it32 and c64 == c32:
            assert np.max(np.abs(l64 - l32) / np.maximum(1.0, np.abs(l64))) < 1e-3
        n_osd += dec64.used_osd
    assert n_osd > 0        # the sample must have exercised OSD
    # zero syndrome -> zero correction after one iteration
    e0, _, it0, c0 = dec64.decode(np.zeros(D, dtype=np.uint8))
    assert e0.sum() == 0 and it0 == 1 and c0


@pytest.mark.parametrize("case", DECODE_CASES)
def test_restated_window_loop_equals_reference_loop(case):
    """oracle/cref.c qo_sw_decode (the C restatement of sliding_window.py:162-186) against the fixtures, which were
    produced by the reference's own Python loop."""
    from oracle import windows as owin
    g = decode_case(case)
    wins = owin.plan(odem.analyze(stimtext.parse_flat(circuit_text(case_circuit(case)))), g["m"], g["W"], g["F"])
    for prec in ("f32", "f64"):
        pred, stats = cref.sw_decode(wins, g["m"], g["K"], g["det"].astype(np.uint8), max_iter=BP_KW["max_iter"],
                                     bp_method=BP_KW["bp_method"], schedule=BP_KW["schedule"], precision=prec)
        assert np.array_equal(pred.astype(np.int64), g["pred_" + prec]), prec


# ---------------------------------------------------------------------------------------------- LSD-0
# A literal, per-cluster restatement of the same published algorithm (python sets and lists, each cluster's local matrix
# solved from scratch after every growth step): pins the C version's incremental machinery (one operation list per cluster,
# survivors keeping their reduction across merges).
def _gf2_greedy_solve(A, s):
    """A [r][c] uint8, s [r]: returns (in_image, x supported on the greedy-first independent columns)."""
    r, c = A.shape
    basis = []   # (reduced vector, pivot row, combination over columns as set)
    piv_cols = []
    for j in range(c):
        v = A[:, j].copy(); comb = {j}
        for (bv, pr, bc) in basis:
            if v[pr]:
                v ^= bv; comb ^= bc
        nz = np.flatnonzero(v)
        if nz.size:
            basis.append((v, int(nz[0]), comb)); piv_cols.append(j)
    z = s.copy(); x = set()
    for (bv, pr, bc) in basis:
        if z[pr]:
            z ^= bv; x ^= bc
    return (not z.any()), sorted(x)

def _lsd0_literal(H, syn, llr, clusters_out=None):
    m, n = H.shape
    rows = [np.flatnonzero(H[i]) for i in range(m)]
    cols = [np.flatnonzero(H[:, j]) for j in range(n)]
    bit_owner = [-1]*n; check_owner=[-1]*m
    cl = {}
    for i in np.flatnonzero(syn):
        cid = len(cl)
        cl[cid] = dict(bits=[], checks=[int(i)], boundary={int(i)}, active=True, valid=False)
        check_owner[i] = cid
    def validate(c):
        A = H[np.ix_(c['checks'], c['bits'])] if c['bits'] else np.zeros((len(c['checks']),0),np.uint8)
        ok, x = _gf2_greedy_solve(A.astype(np.uint8), syn[c['checks']].astype(np.uint8))
        return ok, [c['bits'][k] for k in x]
    inv = sorted(cl)
    while inv:
        for cid in inv:
            c = cl[cid]
            if not c['active']: continue
            cand = []
            for r in sorted(c['boundary']):
                ext = [int(j) for j in rows[r] if bit_owner[j] != cid]
                if not ext: c['boundary'].discard(r)
                cand += ext
            if not cand:
                c['valid'] = True
                continue
            best = min(cand, key=lambda j: (llr[j], j))
            bit_owner[best] = cid; c['bits'].append(best)
            ml = []
            for r in cols[best]:
                o = check_owner[r]
                if o == cid: continue
                if o < 0:
                    check_owner[r] = cid; c['checks'].append(int(r)); c['boundary'].add(int(r))
                elif o not in ml: ml.append(o)
            big = cid
            for o in ml:
                a, b = big, o
                if len(cl[a]['bits']) < len(cl[b]['bits']): small, big = a, b
                else: small, big = b, a
                S, B = cl[small], cl[big]
                for j in S['bits']: bit_owner[j] = big; B['bits'].append(j)
                for r in S['checks']: check_owner[r] = big; B['checks'].append(r)
                B['boundary'] |= S['boundary']
                S['active'] = False
            cl[big]['valid'] = validate(cl[big])[0]
        inv = sorted([k for k in cl if cl[k]['active'] and not cl[k]['valid']], key=lambda k: (len(cl[k]['bits']), k))
    e = np.zeros(n, np.uint8)
    for k in cl:
        if cl[k]['active']:
            for j in validate(cl[k])[1]: e[j] = 1
            if clusters_out is not None:
                clusters_out.append((list(cl[k]['bits']), list(cl[k]['checks'])))
    return e


def _gf2_solve_unique(A, t):
    """x with A x = t for A of full column rank and t in its image (plain Gauss-Jordan; the answer does not depend on pivoting)."""
    A = A.copy().astype(np.uint8); t = t.copy().astype(np.uint8)
    r, c = A.shape
    where = []
    row = 0
    for j in range(c):
        nz = [i for i in range(row, r) if A[i, j]]
        assert nz, "dependent column"
        i = nz[0]
        A[[row, i]] = A[[i, row]]; t[[row, i]] = t[[i, row]]
        for k in range(r):
            if k != row and A[k, j]:
                A[k] ^= A[row]; t[k] ^= t[row]
        where.append(row); row += 1
    assert not t[row:].any(), "outside the image"
    return np.array([t[w] for w in where], dtype=np.uint8)


def _lsdw_literal(H, syn, llr, priors, method, order):
    """LSD beyond order 0 as oracle/cref.c lsd_cluster_higher states it, written per cluster from scratch: local matrix, greedy
    first-independent pivots in column-list order, every candidate solved by a plain elimination, weights summed in list order."""
    import math
    clusters = []
    e = _lsd0_literal(H, syn, llr, clusters)
    for bits, checks in clusters:
        if not bits:
            continue
        A = H[np.ix_(checks, bits)].astype(np.uint8)
        s = syn[checks].astype(np.uint8)
        piv = []                                     # greedy first independent columns
        basis = []
        for k in range(len(bits)):
            v = A[:, k].copy()
            for bv, pr in basis:
                if v[pr]:
                    v ^= bv
            nz = np.flatnonzero(v)
            if nz.size:
                basis.append((v, int(nz[0]))); piv.append(k)
        nonpiv = [k for k in range(len(bits)) if k not in piv]
        wt = [math.log(1.0 / priors[j]) for j in bits]

        def solution(F):
            t = s.copy()
            for k in F:
                t ^= A[:, k]
            x = np.zeros(len(bits), np.uint8)
            x[piv] = _gf2_solve_unique(A[:, piv], t)
            for k in F:
                x[k] = 1
            w = 0.0
            for k in range(len(bits)):
                if x[k]:
                    w += wt[k]
            return w, x
        best_w, best_x = solution([])
        w = min(order, len(nonpiv))
        if method == "lsd_cs":
            cands = [[k] for k in nonpiv] + [[nonpiv[i], nonpiv[j]] for i in range(w) for j in range(i + 1, w)]
        else:
            cands = [[nonpiv[b] for b in range(w) if (pat >> b) & 1] for pat in range(1, 1 << w)]
        for F in cands:
            cw, cx = solution(F)
            if cw < best_w:
                best_w, best_x = cw, cx
        for k, j in enumerate(bits):
            e[j] = best_x[k]
    return e


def test_lsd0_equals_the_literal_restatement():
    rng = np.random.default_rng(5)
    n_lsd = 0
    for trial in range(120):
        m = int(rng.integers(4, 30)); n = int(rng.integers(m, 3 * m + 5))
        H = (rng.random((m, n)) < min(0.5, 3.0 / m)).astype(np.uint8)
        for j in range(n):
            if not H[:, j].any():
                H[rng.integers(m), j] = 1
        p = rng.uniform(0.01, 0.2, n)
        dec = cref.BpOsd(sp.csc_matrix(H), p, max_iter=int(rng.integers(1, 4)), bp_method="minimum_sum", osd_method="lsd_0")
        for t in range(4):
            err = (rng.random(n) < 0.15).astype(np.uint8)
            syn = (H @ err % 2).astype(np.uint8)
            e, llr, it, conv = dec.decode(syn)
            assert np.array_equal(H @ e % 2, syn)
            if not conv:
                n_lsd += 1
                assert np.array_equal(_lsd0_literal(H, syn, llr), e)
    assert n_lsd > 200


@pytest.mark.parametrize("method,order", [("lsd_cs", 1), ("lsd_cs", 4), ("lsd_e", 3)])
def test_lsd_higher_order_equals_the_literal_per_cluster_restatement(method, order):
    """lsd_order > 0 (the reference's phenomenological LSD test asks for order 1, tests/test_decoders.py:124-159): the C
    oracle's incremental version against the literal per-cluster one.  PARITY UNPINNED against ldpc (oracle/cref.c header)."""
    rng = np.random.default_rng(11)
    n_lsd = n_better = 0
    for trial in range(90):
        m = int(rng.integers(4, 30)); n = int(rng.integers(m, 3 * m + 5))
        H = (rng.random((m, n)) < min(0.5, 3.0 / m)).astype(np.uint8)
        for j in range(n):
            if not H[:, j].any():
                H[rng.integers(m), j] = 1
        p = rng.choice([0.02, 0.05, 0.11], n) if trial % 2 else rng.uniform(0.01, 0.2, n)      # few distinct priors: ties everywhere
        kw = dict(max_iter=int(rng.integers(1, 4)), bp_method="minimum_sum")
        dec = cref.BpOsd(sp.csc_matrix(H), p, osd_method=method, osd_order=order, **kw)
        dec0 = cref.BpOsd(sp.csc_matrix(H), p, osd_method="lsd_0", **kw)
        for t in range(4):
            err = (rng.random(n) < 0.15).astype(np.uint8)
            syn = (H @ err % 2).astype(np.uint8)
            e, llr, it, conv = dec.decode(syn)
            assert np.array_equal(H @ e % 2, syn)
            if not conv:
                n_lsd += 1
                assert np.array_equal(_lsdw_literal(H, syn, llr, p, method, order), e)
                e0 = dec0.decode(syn)[0]
                w = lambda x: float(np.log(1.0 / p[x.astype(bool)]).sum())
                assert w(e) <= w(e0) + 1e-9
                n_better += not np.array_equal(e, e0)
    assert n_lsd > 150 and n_better > 5


def test_lsd_exhaustive_order_reaches_the_minimum_weight_inside_every_cluster():
    """Semantic anchor for the higher-order sweep, independent of any elimination detail: with lsd_e and an order that covers all
    non-pivot columns of a cluster, the solution's weight inside the cluster is the minimum over ALL solutions supported on the
    cluster's columns (brute force over 2^bits)."""
    import itertools
    import math
    rng = np.random.default_rng(23)
    checked = 0
    for trial in range(80):
        m = int(rng.integers(4, 16)); n = int(rng.integers(m, 2 * m + 4))
        H = (rng.random((m, n)) < min(0.5, 2.5 / m)).astype(np.uint8)
        for j in range(n):
            if not H[:, j].any():
                H[rng.integers(m), j] = 1
        p = rng.uniform(0.01, 0.3, n)
        dec = cref.BpOsd(sp.csc_matrix(H), p, max_iter=1, bp_method="minimum_sum", osd_method="lsd_e", osd_order=20)
        for t in range(3):
            err = (rng.random(n) < 0.2).astype(np.uint8)
            syn = (H @ err % 2).astype(np.uint8)
            e, llr, it, conv = dec.decode(syn)
            if conv:
                continue
            clusters = []
            _lsd0_literal(H, syn, llr, clusters)
            for bits, checks in clusters:
                if not bits or len(bits) > 14:
                    continue
                A = H[np.ix_(checks, bits)].astype(np.uint8)
                sl = syn[checks].astype(np.uint8)
                wt = np.array([math.log(1.0 / p[j]) for j in bits])
                best = None
                for x in itertools.product((0, 1), repeat=len(bits)):
                    xv = np.array(x, dtype=np.uint8)
                    if np.array_equal(A @ xv % 2, sl):
                        w = float(wt[xv.astype(bool)].sum())
                        best = w if best is None or w < best else best
                assert best is not None
                got = float(wt[e[bits].astype(bool)].sum())
                assert abs(got - best) < 1e-9, (trial, t, bits, got, best)
                checked += 1
    assert checked > 60


def test_lsd0_on_a_decoding_window():
    """On a real window (gross code, p = 3e-3) LSD-0 satisfies the syndrome and stays local: far fewer columns than OSD-0's
    rank-many pivots."""
    from oracle import windows as owin
    case = "bb144_r10_p3e-3_W5F3"
    g = decode_case(case)
    w = owin.plan(odem.analyze(stimtext.parse_flat(circuit_text(case_circuit(case)))), g["m"], g["W"], g["F"])[1]
    H, pri = w["H"], w["priors"]
    dec = cref.BpOsd(H, pri, max_iter=3, bp_method="minimum_sum", osd_method="lsd_0")
    syn = g["det"][:24, w["row0"]:w["row0"] + H.shape[0]].astype(np.uint8)
    used = 0
    for s in syn:
        e, llr, it, conv = dec.decode(s)
        assert np.array_equal(H @ e % 2, s)
        used += dec.used_osd
    assert used >= 3


# ---------------------------------------------------------------------------------------------- OSD, literally
def _osd_literal(H, syn, llr, priors, method, order):
    """OSD as published (Panteleev & Kalachev / Roffe et al.; ldpc's osd.hpp as far as remembered), written the slow obvious way
    on dense numpy arrays: columns by ascending (LLR, index); row reduction with the first row at or below the current rank as
    pivot; OSD-0 = solve on the pivots; osd_cs(w) = every single non-pivot column then every pair among the first w, osd_e(w) =
    every non-empty subset of the first w; a candidate's weight is sum log(1/p_j) over its set bits in column-index order; the
    first strictly lighter candidate wins."""
    m, n = H.shape
    order_idx = sorted(range(n), key=lambda j: (llr[j], j))
    A = np.concatenate([H[:, order_idx], syn.reshape(-1, 1)], axis=1).astype(np.uint8)
    piv, rank = [], 0
    for k in range(n):
        if rank == m:
            break
        nz = np.flatnonzero(A[rank:, k])
        if nz.size == 0:
            continue
        p = rank + int(nz[0])
        A[[rank, p]] = A[[p, rank]]
        for i in range(m):
            if i != rank and A[i, k]:
                A[i] ^= A[rank]
        piv.append(k)
        rank += 1

    def weight(x):
        w = 0.0
        for j in range(n):
            if x[j]:
                w += np.log(1.0 / priors[j])
        return w

    def solution(flips):
        x = np.zeros(n, dtype=np.uint8)
        for r, k in enumerate(piv):
            b = A[r, n]
            for f in flips:
                b ^= A[r, f]
            if b:
                x[order_idx[k]] = 1
        for f in flips:
            x[order_idx[f]] = 1
        return x

    best = solution([])
    if method == "osd_0" or order == 0:
        return best
    bw = weight(best)
    nonpiv = [k for k in range(n) if k not in set(piv)]
    w = min(order, len(nonpiv))
    cands = []
    if method == "osd_cs":
        cands = [[k] for k in nonpiv] + [[nonpiv[i], nonpiv[j]] for i in range(w) for j in range(i + 1, w)]
    else:
        cands = [[nonpiv[b] for b in range(w) if (pat >> b) & 1] for pat in range(1, 1 << w)]
    for fl in cands:
        x = solution(fl)
        cw = weight(x)
        if cw < bw:
            bw, best = cw, x
    return best


@pytest.mark.parametrize("method,order", [("osd_0", 0), ("osd_cs", 1), ("osd_cs", 4), ("osd_e", 3)])
def test_osd_equals_the_literal_restatement(method, order):
    """Pins the C oracle's packed-word elimination, stable sort and candidate sweeps against the obvious dense version,
    including rank-deficient matrices and syndromes outside the image."""
    rng = np.random.default_rng(17 + order)
    n_osd = 0
    for trial in range(60):
        m = int(rng.integers(4, 24)); n = int(rng.integers(m, 3 * m + 4))
        H = (rng.random((m, n)) < min(0.5, 3.0 / m)).astype(np.uint8)
        if trial % 4 == 0 and m > 5:
            H[m - 1] = H[0] ^ H[1]                       # rank deficient
        p = rng.uniform(0.01, 0.2, n)
        dec = cref.BpOsd(sp.csc_matrix(H), p, max_iter=int(rng.integers(1, 3)), bp_method="minimum_sum", osd_method=method, osd_order=order)
        for t in range(3):
            syn = ((H @ (rng.random(n) < 0.2)) % 2).astype(np.uint8) if t else rng.integers(0, 2, m).astype(np.uint8)   # t == 0: any syndrome
            e, llr, it, conv = dec.decode(syn)
            if not conv:
                n_osd += 1
                assert np.array_equal(_osd_literal(H, syn, llr, p, method, order), e), (trial, t)
    assert n_osd > 60


# ---------------------------------------------------------------------------------------------- BP, literally
def _bp_literal(H, syn, priors, max_iter, method, schedule, alpha):
    """ldpc v2's BP as published, on python dicts keyed by (row, column): flooding = all check->bit messages from the previous
    bit->check messages, then all bits; serial = bit by bit in index order from the current messages.  Same floating-point
    operation order as the statement in oracle/bp_impl.inc (prefix sums over a column in ascending row order, then the suffix)."""
    import math
    m, n = H.shape
    rows = [list(np.flatnonzero(H[i])) for i in range(m)]
    cols = [list(np.flatnonzero(H[:, j])) for j in range(n)]
    l0 = [math.log((1.0 - p) / p) for p in priors]
    v = {(i, j): l0[j] for j in range(n) for i in cols[j]}
    llr = list(l0)

    def check_to_bit(i, j, it):
        others = [v[(i, g)] for g in rows[i] if g != j]
        if method == "product_sum":
            t = 1.0
            for x in others:
                t *= math.tanh(x / 2.0)
            sign = -1.0 if syn[i] else 1.0
            num, den = 1.0 + t, 1.0 - t
            if den == 0.0:
                return sign * math.inf
            q = num / den
            return sign * (-math.inf if q == 0.0 else (math.nan if q < 0.0 else math.log(q)))
        a = alpha if alpha != 0.0 else 1.0 - 2.0 ** (-it)
        mag = min([abs(x) for x in others], default=float(np.finfo(np.float64).max))
        sgn = int(syn[i]) + sum(1 for x in others if x <= 0)
        return (a if sgn % 2 == 0 else -a) * mag

    for it in range(1, max_iter + 1):
        e = np.zeros(n, dtype=np.uint8)
        if schedule == "parallel":
            c = {(i, j): check_to_bit(i, j, it) for j in range(n) for i in cols[j]}
        for j in range(n):
            if schedule == "serial":
                c = {(i, j): check_to_bit(i, j, it) for i in cols[j]}
            t = l0[j]
            pre = []
            for i in cols[j]:
                pre.append(t)
                t = t + c[(i, j)]
            llr[j] = t
            e[j] = 1 if t <= 0 else 0
            if schedule == "serial" or True:
                suf = 0.0
                newv = {}
                for k in range(len(cols[j]) - 1, -1, -1):
                    i = cols[j][k]
                    newv[(i, j)] = pre[k] + suf
                    suf = suf + c[(i, j)]
                if schedule == "serial":
                    v.update(newv)
                else:
                    pending = newv if j == 0 else {**pending, **newv}
        if schedule == "parallel":
            v.update(pending)
        if np.array_equal((H @ e) % 2, syn):
            return e, np.array(llr), it, True
    return e, np.array(llr), max_iter, False


@pytest.mark.parametrize("method,schedule,alpha", [("minimum_sum", "parallel", 1.0), ("minimum_sum", "parallel", 0.0), ("minimum_sum", "serial", 0.75),
                                                   ("product_sum", "parallel", 1.0), ("product_sum", "serial", 1.0)])
def test_bp_equals_the_literal_restatement(method, schedule, alpha):
    """Pins the C oracle's edge indexing, sweep order and scaling against the dictionary version: min-sum posteriors bit for bit,
    product-sum to 1e-9 (same libm; the C version multiplies prefix x suffix instead of the others in one run)."""
    rng = np.random.default_rng(41)
    for trial in range(25):
        m = int(rng.integers(4, 16)); n = int(rng.integers(m, 2 * m + 6))
        H = (rng.random((m, n)) < min(0.5, 3.0 / m)).astype(np.uint8)
        for j in range(n):
            if not H[:, j].any():
                H[rng.integers(m), j] = 1
        p = rng.uniform(0.02, 0.2, n)
        dec = cref.BpOsd(sp.csc_matrix(H), p, max_iter=4, bp_method=method, schedule=schedule, ms_scaling_factor=alpha, osd=False)
        for t in range(3):
            syn = ((H @ (rng.random(n) < 0.15)) % 2).astype(np.uint8)
            e, llr, it, conv = dec.decode(syn)
            e2, llr2, it2, conv2 = _bp_literal(H, syn, p, 4, method, schedule, alpha)
            if method == "minimum_sum":
                assert (it, conv) == (it2, conv2) and np.array_equal(llr, llr2) and np.array_equal(e, e2), (trial, t)
            elif it == it2:
                fin = np.isfinite(llr) & np.isfinite(llr2)
                assert np.allclose(llr[fin], llr2[fin], rtol=1e-9, atol=1e-9), (trial, t)


def test_message_free_flooding_min_sum_prototype_equals_the_oracle():
    """tools/proto_stateless_minsum.py: flooding min-sum from (m1, m2, parity) per row and two bits per edge, without a message
    array -- the design DESIGN.md section 7 proposes for windows that do not fit in shared memory -- returns the oracle's
    posteriors, iteration counts and hard decisions bit for bit."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("proto_stateless_minsum", os.path.join(root, "tools", "proto_stateless_minsum.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.self_check(trials=24, seed=7) == 72
