"""BP-OSD sliding-window wrappers; signatures and defaults of reference ``src/quits/decoder/bposd.py:10,54``."""
from __future__ import annotations

from .inner import BpOsdDecoder
from .sliding_window import sliding_window_circuit_mem, sliding_window_phenom_mem


def sliding_window_bposd_phenom_mem(zcheck_samples, hz, lz, W, F, eff_error_rate_per_fault: float = None, max_iter=2, osd_order=0,
                                    bp_method='product_sum', schedule='serial', osd_method='osd_cs', tqdm_on=False,
                                    error_rate: float = None):
    """Drop-in for the reference function of the same name (``src/quits/decoder/bposd.py:10-51``): same arguments, defaults,
    deprecated ``error_rate`` alias and ValueError when neither rate is given."""
    if eff_error_rate_per_fault is None:
        eff_error_rate_per_fault = error_rate
    if eff_error_rate_per_fault is None:
        raise ValueError("eff_error_rate_per_fault must be provided (or use deprecated error_rate).")
    params = lambda: {'bp_method': bp_method, 'max_iter': max_iter, 'schedule': schedule, 'osd_method': osd_method,
                      'osd_order': osd_order, 'error_rate': float(eff_error_rate_per_fault)}
    return sliding_window_phenom_mem(zcheck_samples, hz, lz, W, F, BpOsdDecoder, BpOsdDecoder, params(), params(), 'decode', 'decode',
                                     tqdm_on=tqdm_on)


def sliding_window_bposd_circuit_mem(zcheck_samples, circuit, hz, lz, W, F, max_iter=2, osd_order=0, bp_method='product_sum',
                                     schedule='serial', osd_method='osd_cs', tqdm_on=False):
    """Drop-in for the reference function of the same name, same defaults.  The GPU kernels cover ``bp_method`` 'minimum_sum'
    and 'product_sum', ``schedule`` 'parallel' and 'serial', and ``osd_method`` 'osd_0' / 'osd_cs' (order <= 32) / 'osd_e'
    (order <= 12); anything else raises NotImplementedError -- there is no CPU fallback."""
    params = lambda: {'bp_method': bp_method, 'max_iter': max_iter, 'schedule': schedule, 'osd_method': osd_method,
                      'osd_order': osd_order}
    return sliding_window_circuit_mem(zcheck_samples, circuit, hz, lz, W, F, BpOsdDecoder, BpOsdDecoder, params(), params(),
                                      'channel_probs', 'channel_probs', 'decode', 'decode', tqdm_on=tqdm_on)


__all__ = ["sliding_window_bposd_phenom_mem", "sliding_window_bposd_circuit_mem"]
