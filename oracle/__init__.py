"""TEST INFRASTRUCTURE ONLY — CPU oracle for the OptionsLab Monte Carlo hot path.

Nothing in ``optionslab_b200`` imports this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the timed CPU baseline.

* ``oracle.reference_mc``  — NumPy restatement of the reference's arithmetic
  (each function cites the reference file:line it follows).  Parity is PINNED:
  ``tests/golden/reference_goldens.json`` was produced by importing the real
  reference from ``/root/reference`` (script: ``tests/golden/make_goldens.py``)
  and ``tests/test_oracle_golden.py`` checks the restatement against it bit for
  bit (same NumPy) / to 1e-13 (other NumPy builds).
* ``oracle/build_ref.py``   — installs the UNMODIFIED reference package into ``oracle/_ref/`` (git-ignored build
  output, shipped to the GPU box by gpurun) so that ``bench.py --impl reference`` / ``cpu_baseline`` time the
  reference's own code and ``tests/test_reference_install.py`` checks the restatement against it bit for bit.
* ``oracle/philox_oracle.c`` — plain-C restatement of Philox4x32-10 (Salmon et
  al., SC'11; Random123 v1.14 ``philox.h``) and of this repo's documented
  uniform→normal mapping.  The reference pins nothing about Philox (its RNG is
  NumPy's PCG64 / MT19937), so the Philox stream itself is pinned only by the
  Random123 known-answer vectors: "parity unpinned" with respect to the
  reference, by construction.
"""
