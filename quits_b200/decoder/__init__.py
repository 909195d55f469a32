"""Drop-in for ``quits.decoder`` (reference ``src/quits/decoder/__init__.py:13-24``): same exported names."""
from .base import WindowPlan, detector_error_model_to_matrix, spacetime
from .bplsd import sliding_window_bplsd_circuit_mem, sliding_window_bplsd_phenom_mem
from .bposd import sliding_window_bposd_circuit_mem, sliding_window_bposd_phenom_mem
from .inner import BpLsdDecoder, BpOsdDecoder
from .sliding_window import sliding_window_circuit_mem, sliding_window_phenom_mem

__all__ = [
    "detector_error_model_to_matrix", "spacetime",
    "sliding_window_phenom_mem", "sliding_window_circuit_mem",
    "sliding_window_bposd_phenom_mem", "sliding_window_bposd_circuit_mem",
    "sliding_window_bplsd_phenom_mem", "sliding_window_bplsd_circuit_mem",
    "BpOsdDecoder", "BpLsdDecoder", "WindowPlan",
]
