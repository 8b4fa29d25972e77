"""CPU oracle for the MiCo hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a plain-PyTorch fp32 CPU restatement of the
reference algorithm (``/root/reference`` = invictus717/MiCo @ 831847f), written in
functional style over the reference's own ``state_dict`` keys.  It exists to
*check* the CUDA product in ``mico_b200/`` and to serve as the timed CPU arm of
``bench.py`` (``cpu_baseline`` / ``--impl reference``).

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
    import this package;
  * nothing under ``mico_b200/`` may import, call or fall back to it.

Parity status: the reference holds NO golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference itself,
imported in the build container through ``oracle/ref_shims.py`` by
``oracle/make_golden.py``; the resulting fixtures live in ``tests/golden/`` and are
re-checked on every CPU test run (tests/test_oracle_golden.py).  Library versions
at pin time are recorded in ``tests/golden/MANIFEST.json`` ("parity pinned to the
reference as run under torch 2.11 / transformers 5.5 in this image").
"""
