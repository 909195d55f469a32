"""DEM -> check matrix conversion and sliding-window slicing, host side (set-up, once per circuit).

Mirrors the reference interface ``detector_error_model_to_matrix`` / ``spacetime``
(reference ``src/quits/decoder/base.py:74-127,134-190``): same names, arguments, return values and
exceptions.  The work itself is done by the C++ host code behind ``qb_dem_*`` / ``qb_plan_*``
(quits_b200/csrc/qb_host.cpp), which is also what feeds the CUDA decoder, so what these functions return
is exactly what the GPU decodes with.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.sparse import csc_matrix

from .. import _native as N
from ..circuit import Circuit, DetectorErrorModel


def _as_dem(dem) -> DetectorErrorModel:
    """Our DEM as is; a foreign (stim-shaped) DEM is re-read through the calls the reference itself makes on it
    (``flattened()``, ``type``, ``args_copy()``, ``targets_copy()``, ``is_relative_detector_id()`` ..., base.py:101-125)."""
    if isinstance(dem, DetectorErrorModel):
        return dem
    probs, dptr, didx, optr, oidx = [], [0], [], [0], []
    for ins in dem.flattened():
        if ins.type == "error":
            probs.append(ins.args_copy()[0])
            for t in ins.targets_copy():
                if t.is_relative_detector_id():
                    didx.append(t.val)
                elif t.is_logical_observable_id():
                    oidx.append(t.val)
            dptr.append(len(didx))
            optr.append(len(oidx))
        elif ins.type in ("detector", "logical_observable"):
            pass
        else:
            raise NotImplementedError()
    out = DetectorErrorModel.__new__(DetectorErrorModel)
    h = C.c_void_p()
    a = [np.ascontiguousarray(probs, dtype=np.float64), np.ascontiguousarray(dptr, dtype=np.int64),
         np.ascontiguousarray(didx if didx else [0], dtype=np.int32), np.ascontiguousarray(optr, dtype=np.int64),
         np.ascontiguousarray(oidx if oidx else [0], dtype=np.int32)]
    N.check(N.lib().qb_dem_from_errors(int(dem.num_detectors), int(dem.num_observables), len(probs), *[N.ptr(x) for x in a],
                                       C.byref(h)))
    out._h = h
    sizes = np.zeros(9, dtype=np.int64)
    N.check(N.lib().qb_dem_sizes(h, N.ptr(sizes)))
    (out.num_detectors, out.num_observables, out.num_errors, out._nnz_det, out._nnz_obs, out.num_columns, out._nnz_h,
     out._nnz_l, out.num_detectorless) = (int(x) for x in sizes)
    return out


def _csc(ptr, idx, shape):
    return csc_matrix((np.ones(len(idx), dtype=np.uint8), np.asarray(idx, dtype=np.int32), np.asarray(ptr, dtype=np.int32)),
                      shape=shape)


def detector_error_model_to_matrix(dem):
    """(check_matrix csc uint8 [D x C], observables_matrix csc [K x C], priors float64 [C]); reference base.py:74-127."""
    d = _as_dem(dem)
    if d.num_detectorless:
        # the reference prints every detector-less error instruction (base.py:114-115)
        for ins in d:
            if not any(t.is_relative_detector_id() for t in ins.targets_copy()):
                print(ins)
    h_ptr, h_idx, l_ptr, l_idx, priors = d.matrix_arrays()
    return (_csc(h_ptr, h_idx, (d.num_detectors, d.num_columns)), _csc(l_ptr, l_idx, (d.num_observables, d.num_columns)), priors)


class WindowPlan:
    """Host-side window plan (``qb_plan``): what ``spacetime`` returns, in the arrays the GPU decoder consumes."""

    def __init__(self, dem, m: int, W: int, F: int, num_cor_rounds: int = -1):
        self.dem = _as_dem(dem)
        h = C.c_void_p()
        N.check(N.lib().qb_plan_create(self.dem._h, int(m), int(W), int(F), int(num_cor_rounds), C.byref(h)))
        self._h = h
        info = np.zeros(8, dtype=np.int64)
        N.check(N.lib().qb_plan_info(self._h, N.ptr(info)))
        (self.n_windows, self.m, self.K, self.D, self.W, self.F, self.num_rounds, wh) = (int(x) for x in info)
        self.whole_history = bool(wh)

    @classmethod
    def explicit(cls, m: int, K: int, D: int, windows) -> "WindowPlan":
        """Plan from explicit windows (``qb_plan_create_explicit``).  Each window is a dict with ``row0``, ``H`` (scipy
        sparse, rows x ncols), ``priors`` (ncols), ``L`` (K x ncommit) and ``U`` (m x ncommit, or None for the last window);
        ``ncommit`` is the column count of ``L``."""
        dims, hp, hi, pr, lp, li, up, ui = [], [], [], [], [], [], [], []
        for k, w in enumerate(windows):
            H = csc_matrix(w["H"]); H.sort_indices()
            L = csc_matrix(w["L"]); L.sort_indices()
            ncommit = L.shape[1]
            U = w.get("U")
            if U is None:
                U = csc_matrix((0, ncommit), dtype=np.uint8)
            U = csc_matrix(U); U.sort_indices()
            dims.append([int(w["row0"]), H.shape[0], 0, H.shape[1], ncommit, H.nnz, L.nnz, U.nnz, int(w.get("urow0", 0)), U.shape[0]])
            hp.append(H.indptr.astype(np.int64)); hi.append(H.indices.astype(np.int32))
            pr.append(np.ascontiguousarray(w["priors"], dtype=np.float64))
            lp.append(L.indptr.astype(np.int64)); li.append(L.indices.astype(np.int32))
            up.append(U.indptr.astype(np.int64)); ui.append(U.indices.astype(np.int32))
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if sum(len(x) for x in xs) else np.zeros(1, dtype=dt), dtype=dt)
        dims = np.ascontiguousarray(dims, dtype=np.int64)
        self = cls.__new__(cls)
        self.dem = None
        h = C.c_void_p()
        N.check(N.lib().qb_plan_create_explicit(int(m), int(K), int(D), len(windows), N.ptr(dims), N.ptr(cat(hp, np.int64)),
                                                N.ptr(cat(hi, np.int32)), N.ptr(cat(pr, np.float64)), N.ptr(cat(lp, np.int64)),
                                                N.ptr(cat(li, np.int32)), N.ptr(cat(up, np.int64)), N.ptr(cat(ui, np.int32)), C.byref(h)))
        self._h = h
        info = np.zeros(8, dtype=np.int64)
        N.check(N.lib().qb_plan_info(self._h, N.ptr(info)))
        (self.n_windows, self.m, self.K, self.D, self.W, self.F, self.num_rounds, wh) = (int(x) for x in info)
        self.whole_history = bool(wh)
        return self

    def __del__(self):
        try:
            if getattr(self, "_h", None) and N.alive():
                N.lib().qb_plan_free(self._h)
                self._h = None
        except Exception:
            pass

    def window(self, k: int) -> dict:
        dims = np.zeros(10, dtype=np.int64)
        N.check(N.lib().qb_plan_window(self._h, int(k), N.ptr(dims), *([None] * 7)))
        row0, rows, col0, ncols, ncommit, nnz, nnz_l, nnz_u, urow0, urows = (int(x) for x in dims)
        h_ptr = np.zeros(ncols + 1, dtype=np.int64)
        h_idx = np.zeros(max(nnz, 1), dtype=np.int32)
        priors = np.zeros(max(ncols, 1), dtype=np.float64)
        l_ptr = np.zeros(ncommit + 1, dtype=np.int64)
        l_idx = np.zeros(max(nnz_l, 1), dtype=np.int32)
        u_ptr = np.zeros(ncommit + 1, dtype=np.int64)
        u_idx = np.zeros(max(nnz_u, 1), dtype=np.int32)
        N.check(N.lib().qb_plan_window(self._h, int(k), None, N.ptr(h_ptr), N.ptr(h_idx), N.ptr(priors), N.ptr(l_ptr), N.ptr(l_idx),
                                       N.ptr(u_ptr), N.ptr(u_idx)))
        return {"row0": row0, "rows": rows, "col0": col0, "ncols": ncols, "ncommit": ncommit, "urow0": urow0, "urows": urows,
                "H": _csc(h_ptr, h_idx[:nnz], (rows, ncols)), "priors": priors[:ncols],
                "L": _csc(l_ptr, l_idx[:nnz_l], (self.K, ncommit)),
                "U": _csc(u_ptr, u_idx[:nnz_u], (urows, ncommit)) if urows else None}


def spacetime(circuit, hz, W, F, num_cor_rounds):
    """(window_check_set, window_observable_set, window_priors_set, window_update); reference base.py:134-190."""
    if F == 0:
        raise ValueError("Input parameter F cannot be zero.")
    model = Circuit.of(circuit).detector_error_model(decompose_errors=False)
    plan = WindowPlan(model, hz.shape[0], W, F, num_cor_rounds)
    checks, observables, priors, updates = [], [], [], []
    for k in range(plan.n_windows):
        w = plan.window(k)
        checks.append(w["H"])
        observables.append(w["L"])
        priors.append(w["priors"])
        if k < plan.n_windows - 1:
            updates.append(w["U"])
    return checks, observables, priors, updates


__all__ = ["detector_error_model_to_matrix", "spacetime", "WindowPlan"]
