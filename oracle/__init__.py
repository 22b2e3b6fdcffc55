"""CPU oracle for the ACE SFNO inference hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a CPU (numpy / torch-CPU) restatement of the reference algorithm
(ai2cm/ace ``fme`` 2026.5.1 + the init-time routines of torch-harmonics 0.8.0)
for exactly one path: ``SphericalFourierNeuralOperatorNet.forward`` and the
``RealSHT`` / ``InverseRealSHT`` transforms underneath it.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under
``ace_b200/`` imports it; the product path raises if the CUDA library is
missing instead of falling back to this code.

Parity status: PINNED.  ``oracle/make_golden.py`` (run in the build container,
where ``/root/reference`` exists) checks every function here against
  * the reference's own stored goldens
      fme/core/benchmark/testdata/sht-regression.pt
      fme/core/benchmark/testdata/inverse_sht-regression.pt
      fme/ace/models/modulus/testdata/test_sfnonet_output_is_unchanged.pt
  * the live reference modules imported from /root/reference (``oracle/refload.py``)
and writes the vectors it used to ``tests/golden/``; ``tests/test_oracle_*.py``
re-check the oracle against those committed vectors on every run.
"""
