"""BP-LSD sliding-window wrappers; signatures and defaults of reference ``src/quits/decoder/bplsd.py:10,54``."""
from __future__ import annotations

from .inner import BpLsdDecoder
from .sliding_window import sliding_window_circuit_mem, sliding_window_phenom_mem


def sliding_window_bplsd_phenom_mem(zcheck_samples, hz, lz, W, F, eff_error_rate_per_fault: float = None, max_iter=2, lsd_order=0,
                                    bp_method='product_sum', schedule='serial', lsd_method='lsd_cs', tqdm_on=False,
                                    error_rate: float = None):
    """Drop-in for the reference function of the same name (``src/quits/decoder/bplsd.py:10-51``): same arguments, defaults,
    deprecated ``error_rate`` alias and ValueError when neither rate is given."""
    if eff_error_rate_per_fault is None:
        eff_error_rate_per_fault = error_rate
    if eff_error_rate_per_fault is None:
        raise ValueError("eff_error_rate_per_fault must be provided (or use deprecated error_rate).")
    params = lambda: {'bp_method': bp_method, 'max_iter': max_iter, 'schedule': schedule, 'lsd_method': lsd_method,
                      'lsd_order': lsd_order, 'error_rate': float(eff_error_rate_per_fault)}
    return sliding_window_phenom_mem(zcheck_samples, hz, lz, W, F, BpLsdDecoder, BpLsdDecoder, params(), params(), 'decode', 'decode',
                                     tqdm_on=tqdm_on)


def sliding_window_bplsd_circuit_mem(zcheck_samples, circuit, hz, lz, W, F, max_iter=2, lsd_order=0, bp_method='product_sum',
                                     schedule='serial', lsd_method='lsd_cs', tqdm_on=False):
    params = lambda: {'bp_method': bp_method, 'max_iter': max_iter, 'schedule': schedule, 'lsd_method': lsd_method,
                      'lsd_order': lsd_order}
    return sliding_window_circuit_mem(zcheck_samples, circuit, hz, lz, W, F, BpLsdDecoder, BpLsdDecoder, params(), params(),
                                      'channel_probs', 'channel_probs', 'decode', 'decode', tqdm_on=tqdm_on)


__all__ = ["sliding_window_bplsd_phenom_mem", "sliding_window_bplsd_circuit_mem"]
