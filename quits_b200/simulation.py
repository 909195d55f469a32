"""Drop-in for ``quits.simulation.get_stim_mem_result`` (reference ``src/quits/simulation.py:8-28``)."""
from __future__ import annotations

import numpy as np

from .circuit import Circuit


def get_stim_mem_result(circuit, num_trials, seed=-1):
    """Pauli-frame Monte-Carlo of the memory circuit on the GPU (K1 kernel).

    :param circuit: Stim circuit -- a ``quits_b200.Circuit``, Stim text, or any object whose ``str()`` is Stim text
                    (a ``stim.Circuit`` is)
    :param num_trials: number of shots
    :param seed: run seed; negative => drawn from OS entropy, as the reference does when no seed is given
    :return: (detection_events bool[num_trials, #detectors], observable_flips bool[num_trials, #observables])
    """
    c = Circuit.of(circuit)
    if seed is None or seed < 0:
        seed = int(np.random.SeedSequence().generate_state(2, dtype=np.uint32).view(np.uint64)[0]) & (2**63 - 1)
    return c.sample(int(num_trials), int(seed))


__all__ = ["get_stim_mem_result"]
