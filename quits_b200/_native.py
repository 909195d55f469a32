"""ctypes binding of libquits_b200.so (include/quits_b200.h).  There is no CPU fallback: if the library is
missing the import fails loudly, and every compute entry point fails loudly without a CUDA device."""
from __future__ import annotations

import atexit
import ctypes as C
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("QB_LIB") or os.path.join(_HERE, "libquits_b200.so")      # QB_LIB: A/B builds of the same ABI (development)


class QbStats(C.Structure):
    _fields_ = [("shots", C.c_int64), ("windows", C.c_int64), ("bp_converged", C.c_int64), ("bp_iterations", C.c_int64),
                ("osd_calls", C.c_int64), ("bp_launches", C.c_int64), ("osd_launches", C.c_int64),
                ("frame_launches", C.c_int64), ("other_launches", C.c_int64),
                ("frame_ms", C.c_double), ("bp_ms", C.c_double), ("osd_ms", C.c_double), ("total_ms", C.c_double),
                ("bp_alg_bytes", C.c_double), ("osd_alg_bytes", C.c_double), ("frame_alg_bytes", C.c_double),
                ("osd_columns", C.c_int64), ("osd_pivots", C.c_int64), ("osd_max_columns", C.c_int64), ("osd_overflows", C.c_int64),
                ("bp_edge_iters", C.c_double)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class QbBpOpts(C.Structure):
    _fields_ = [("bp_method", C.c_int32), ("schedule", C.c_int32), ("max_iter", C.c_int32), ("ms_scaling_factor", C.c_double),
                ("osd_method", C.c_int32), ("osd_order", C.c_int32), ("precision", C.c_int32), ("capacity", C.c_int32),
                ("profile", C.c_int32), ("lanes", C.c_int32)]


class QbCircuitInfo(C.Structure):
    _fields_ = [("n_qubits", C.c_int32), ("n_measurements", C.c_int32), ("n_detectors", C.c_int32), ("n_observables", C.c_int32),
                ("n_flat_ops", C.c_int64), ("n_tape_ops", C.c_int64), ("n_noise_sites", C.c_int64), ("ring", C.c_int32)]


# every symbol include/quits_b200.h declares (tests check the library exports all of them)
SYMBOLS = ["qb_last_error", "qb_version", "qb_build_info", "qb_device_count", "qb_host_alloc", "qb_host_free", "qb_ctx_create", "qb_ctx_destroy", "qb_ctx_synchronize",
           "qb_circuit_parse", "qb_circuit_free", "qb_circuit_get_info", "qb_circuit_flat", "qb_sample", "qb_sample_packed",
           "qb_sample_faults", "qb_dem_from_circuit", "qb_dem_free", "qb_dem_sizes", "qb_dem_errors", "qb_dem_matrix",
           "qb_dem_from_errors", "qb_plan_create", "qb_plan_create_explicit", "qb_plan_free", "qb_plan_info", "qb_plan_window", "qb_plan_layout",
           "qb_sw_create", "qb_sw_create_single", "qb_sw_free", "qb_sw_decode",
           "qb_sw_decode_packed", "qb_bp_decode_batch", "qb_mc_run"]

_lib = None

# Set at interpreter exit: from then on no destructor calls into the CUDA runtime (the driver may already be unloading, and the
# order in which Python drops the remaining objects is arbitrary); the OS reclaims everything.
_alive = True


def _shutdown():
    global _alive
    _alive = False


atexit.register(_shutdown)


def alive() -> bool:
    return _alive


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            "quits_b200: %s is missing. Build it with `python -m quits_b200.build` (needs nvcc); "
            "there is no CPU fallback." % SO_PATH)
    L = C.CDLL(SO_PATH)
    stale = [name for name in SYMBOLS if not hasattr(L, name)]
    if stale:
        raise ImportError("quits_b200: %s does not export %s -- it was built from older sources. Rebuild it with "
                          "`python -m quits_b200.build`; there is no CPU fallback." % (SO_PATH, ", ".join(stale)))
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    L.qb_last_error.restype = C.c_char_p
    L.qb_version.restype = C.c_int
    L.qb_build_info.restype = C.c_char_p
    L.qb_device_count.restype = C.c_int
    L.qb_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.qb_host_free.argtypes = [vp]
    L.qb_host_free.restype = None
    L.qb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.qb_ctx_destroy.argtypes = [vp]
    L.qb_ctx_destroy.restype = None
    L.qb_ctx_synchronize.argtypes = [vp]
    L.qb_circuit_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.qb_circuit_free.argtypes = [vp]
    L.qb_circuit_free.restype = None
    L.qb_circuit_get_info.argtypes = [vp, C.POINTER(QbCircuitInfo)]
    L.qb_circuit_flat.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i64)]
    L.qb_sample.argtypes = [vp, vp, u64, u64, u64, vp, vp]
    L.qb_sample_packed.argtypes = [vp, vp, u64, u64, u64, vp, vp]
    L.qb_sample_faults.argtypes = [vp, vp, i64, vp, vp, vp, vp, u64, vp, vp]
    L.qb_dem_from_circuit.argtypes = [vp, C.POINTER(vp)]
    L.qb_dem_free.argtypes = [vp]
    L.qb_dem_free.restype = None
    L.qb_dem_sizes.argtypes = [vp, vp]
    L.qb_dem_errors.argtypes = [vp] + [vp] * 8
    L.qb_dem_matrix.argtypes = [vp] + [vp] * 5
    L.qb_dem_from_errors.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    L.qb_plan_create.argtypes = [vp, i32, i32, i32, i32, C.POINTER(vp)]
    L.qb_plan_create_explicit.argtypes = [i32, i32, i32, i32] + [vp] * 8 + [C.POINTER(vp)]
    L.qb_plan_free.argtypes = [vp]
    L.qb_plan_free.restype = None
    L.qb_plan_info.argtypes = [vp, vp]
    L.qb_plan_window.argtypes = [vp, i32, vp] + [vp] * 7
    L.qb_plan_layout.argtypes = [vp, i32, i32, vp, vp, vp]
    L.qb_sw_create.argtypes = [vp, vp, C.POINTER(QbBpOpts), C.POINTER(vp)]
    L.qb_sw_create_single.argtypes = [vp, i32, i32, vp, vp, vp, C.POINTER(QbBpOpts), C.POINTER(vp)]
    L.qb_sw_free.argtypes = [vp]
    L.qb_sw_free.restype = None
    L.qb_sw_decode.argtypes = [vp, vp, u64, vp, C.POINTER(QbStats)]
    L.qb_sw_decode_packed.argtypes = [vp, vp, u64, vp, C.POINTER(QbStats)]
    L.qb_bp_decode_batch.argtypes = [vp, vp, u64, vp, vp, vp, vp]
    L.qb_mc_run.argtypes = [vp, vp, vp, u64, u64, u64, vp, C.POINTER(QbStats)]
    _lib = L
    return L


class QbCudaError(RuntimeError):
    pass


def check(rc: int) -> None:
    """Map a status code to the exception type the reference would raise at that point."""
    if rc == 0:
        return
    msg = lib().qb_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    if rc == 2:
        raise NotImplementedError(msg)
    if rc == 4:
        raise TypeError(msg)
    raise QbCudaError(msg)


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def device_count() -> int:
    return int(lib().qb_device_count())


class _PinnedPool:
    """Recycles page-locked host buffers behind numpy arrays.  cudaHostAlloc costs milliseconds per 100 MB, so buffers are
    kept (rounded up to a power of two, at most ``limit`` bytes idle) and handed out again once the last numpy view of
    them is gone."""

    def __init__(self, limit=4 << 30):
        self.free = {}
        self.idle = 0
        self.limit = limit
        self.lock = threading.Lock()

    def _release(self, ptr, cap):
        if not _alive:
            return
        with self.lock:
            if self.idle + cap <= self.limit:
                self.free.setdefault(cap, []).append(ptr)
                self.idle += cap
                return
        lib().qb_host_free(C.c_void_p(ptr))

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        if nbytes < (1 << 20) or nbytes > (2 << 30):  # small arrays, and arrays too large to page-lock sensibly: ordinary memory
            return np.empty(shape, dtype=dtype)
        cap = 1 << (nbytes - 1).bit_length()
        with self.lock:
            lst = self.free.get(cap)
            ptr = lst.pop() if lst else None
            if ptr is not None:
                self.idle -= cap
        if ptr is None:
            p = C.c_void_p()
            if lib().qb_host_alloc(cap, C.byref(p)) != 0 or not p.value:
                return np.empty(shape, dtype=dtype)    # page-locking failed (limits, fragmentation): pageable memory still works
            ptr = p.value
        raw = (C.c_uint8 * nbytes).from_address(ptr)
        weakref.finalize(raw, self._release, ptr, cap).atexit = False
        return np.frombuffer(raw, dtype=dtype).reshape(shape)


_pool = _PinnedPool()


def empty(shape, dtype):
    """Uninitialised numpy array for results that come back from the device (page-locked when large)."""
    return _pool.empty(shape, dtype)
