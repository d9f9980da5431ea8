"""CPU oracle for the line bundle-adjustment hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker or as the timed CPU baseline.  The product path
(``amuya_sfm_lines_beta_b200``) never imports this package and has no CPU
fallback.

Parity pinning: the reference repository ships no tests,
golden vectors or recorded outputs for this path, so the restatement is pinned
against *outputs of the reference itself*, generated in the build container by
importing ``/root/reference/src/bundle_adjustment.py`` (see
``tests/golden/make_golden.py``) and committed as ``tests/golden/*.npz``.
The solver that the reference delegates to (scipy ``least_squares`` TRF + LSMR
with a finite-difference Jacobian) is third-party code that is *replaced*, not
restated; the LM / Schur / PCG algorithm in ``line_ba_oracle.py`` restates the
algorithm of this repository, and trajectory parity is device-vs-oracle.
"""
