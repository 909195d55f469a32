"""CPU oracle for the QUITS Monte-Carlo hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``quits_b200/`` may import this package.  It is imported by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` only, and there only as the checker
(or as the CPU arm being timed), never as the product path.

Parity status: **parity unpinned** at the stim/ldpc boundary.  The reference
(mkangquantum/quits) delegates the arithmetic of this path to the PyPI wheels
``stim>=1.13.0`` and ``ldpc>=2.1.2`` (reference ``pyproject.toml:29-36``), neither
of which is vendored, installed here, or installable (no network).  The
reference's own tests hold no bit-exact vector for the path
(``tests/test_sliding_window.py:102-103`` asserts ``pL <= 0.2`` on 50 shots).  The
oracle therefore restates the published algorithms of those wheels
(Pauli-frame propagation, backward error analysis, BP min-sum/product-sum
flooding, OSD-0) and is pinned three ways:

* the only bit-exact artefact the reference holds for this path, the printed
  [[72,12,6]] circuit of ``doc/02A_custom_circuit_generation.ipynb:82-289``, is
  reproduced through the oracle's parser/printer (tests/test_oracle.py::test_doc02A_printout_is_reproduced);
* forward frame propagation of every detector-error-model column's
  representative fault reproduces that column (frame sim <-> DEM analyser
  self-consistency, tests/test_oracle.py::test_forward_frame_reproduces_every_dem_column);
* the *unmodified* reference window glue (``decoder/base.py:74-190``,
  ``decoder/sliding_window.py:104-188``) is run in the build container on top
  of the oracle through stim/ldpc-shaped shims (``oracle/shims.py``); its
  outputs are the committed fixtures under ``tests/golden/`` (generator:
  ``tools/make_golden.py``).
"""
