"""quits_b200 -- B200-native engine for the QUITS Monte-Carlo hot path.

Pauli-frame sampling of the syndrome-extraction circuit and sliding-window BP(+OSD) decoding over the
circuit-level detector error matrix, as hand-written sm_100a CUDA behind a C ABI (include/quits_b200.h).
The Python layer mirrors the reference's interface for this path: ``get_stim_mem_result``
(reference src/quits/simulation.py:8) and ``quits.decoder`` (src/quits/decoder/__init__.py:13-24).
Importing this package loads libquits_b200.so; there is no CPU fallback.
"""
import sys as _sys

from . import _native

# `python -m quits_b200.build` imports this package before it runs: the one entry that must work while the library is missing or stale
_BUILDING = "quits_b200.build" in getattr(_sys, "orig_argv", ())

if not _BUILDING:
    _native.lib()          # fail loudly at import time if the CUDA library is missing

    from .circuit import Circuit, Context, DetectorErrorModel  # noqa: E402
    from .decoder import (BpLsdDecoder, BpOsdDecoder, detector_error_model_to_matrix, sliding_window_bplsd_circuit_mem,  # noqa: E402
                          sliding_window_bplsd_phenom_mem, sliding_window_bposd_circuit_mem, sliding_window_bposd_phenom_mem,
                          sliding_window_circuit_mem, sliding_window_phenom_mem, spacetime)
    from .decoder.sliding_window import clear_decoder_cache  # noqa: E402
    from .devices import active_devices, set_devices  # noqa: E402
    from .engine import MonteCarlo, SlidingWindowDecoder, run_sharded, shard_range  # noqa: E402
    from .simulation import count_logical_errors, get_codecap_pL, get_stim_mem_result  # noqa: E402

__version__ = "0.1.0"


def build_info() -> str:
    """How the loaded libquits_b200.so was built (`qb_build_info`): ABI, arch, nvcc version, flags and the hash of the sources it
    was compiled from -- equal to ``quits_b200.build.source_hash()`` when the library belongs to this source tree."""
    return _native.lib().qb_build_info().decode("ascii", "replace")
