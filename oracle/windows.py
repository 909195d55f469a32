"""Oracle: DEM -> check matrix -> sliding-window slices, restated with scipy.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates reference ``src/quits/decoder/base.py:74-127`` (``detector_error_model_to_matrix``: key = detector set, repeated
keys XOR-combine the probability, first sighting's observables win, columns numbered by first sighting) and
``base.py:134-190`` / ``sliding_window.py:130-141`` (window count, row/column ranges, committed prefix, carry block).
Pinned by tests/test_oracle.py against tests/golden/windows/*.json, which hold the digests of what the reference's own
``spacetime()`` returned for the same circuits (tools/make_golden.py).  The GPU box has no reference tree, so the CPU
arms of bench.py and the live GPU-vs-oracle tests get their windows from here.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def dem_matrices(dem):
    """oracle.dem.Dem -> (H csc uint8 [D x C], L csc uint8 [K x C], priors float64 [C])."""
    index, cols, obs_of, priors = {}, [], [], []
    for p, dets, obs in zip(dem.probs, dem.dets, dem.obs):
        key = frozenset(dets)
        j = index.get(key)
        if j is None:
            index[key] = len(cols)
            cols.append(sorted(key))
            obs_of.append(sorted(set(obs)))
            priors.append(p)
        else:
            priors[j] = priors[j] * (1 - p) + p * (1 - priors[j])
    C = len(cols)

    def csc(lists, nrows):
        ptr = np.zeros(C + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([len(x) for x in lists])
        idx = np.array([v for x in lists for v in x], dtype=np.int32)
        return sp.csc_matrix((np.ones(len(idx), dtype=np.uint8), idx, ptr), shape=(nrows, C))

    return csc(cols, dem.n_det), csc(obs_of, dem.n_obs), np.array(priors, dtype=np.float64)


def num_windows(D, m, W, F):
    rounds = D // m - 2
    if 2 + rounds - W >= 0:
        n = (2 + rounds - W) // F
        if (2 + rounds - W) % F:
            n += 1
        return n
    return 0


def plan(dem, m, W, F):
    """List of window dicts {row0, H, priors, L, U} in decoding order (U is None for the last window)."""
    if F == 0:
        raise ValueError("Input parameter F cannot be zero.")
    H, L, priors = dem_matrices(dem)
    n_cor = num_windows(dem.n_det, m, W, F)
    out, col_min = [], 0
    for k in range(n_cor):
        sub = H[k * F * m:(k * F + W) * m, col_min:]
        touched = np.flatnonzero(np.diff(sub.indptr) > 0)
        if sub.shape[1] == 0 or touched.size == 0:
            raise ValueError("a decoding window is not touched by any fault")
        sub = sub[:, :touched.max() + 1]
        first = np.flatnonzero(np.diff(sub[:F * m, :].indptr) > 0)
        ncommit = int(first.max()) + 1
        out.append({"row0": k * F * m, "H": sub.tocsc(), "priors": priors[col_min:col_min + sub.shape[1]],
                    "L": L[:, col_min:col_min + ncommit].tocsc(),
                    "U": H[(k + 1) * F * m:((k + 1) * F + 1) * m, col_min:col_min + ncommit].tocsc()})
        col_min += ncommit
    out.append({"row0": F * n_cor * m, "H": H[F * n_cor * m:, col_min:].tocsc(), "priors": priors[col_min:],
                "L": L[:, col_min:].tocsc(), "U": None})
    return out
