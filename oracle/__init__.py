"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the MRLA hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
(``mrla_b200``) never imports this package and has no CPU fallback.

Parity status: the reference (joyfang1106/MRLA) ships no tests, golden vectors or
fixtures of its own ("parity unpinned by the reference").  This oracle is instead
pinned against the *reference itself*: ``tests/golden/make_golden.py`` imports the
unmodified reference classes from ``/root/reference`` in the dev container, runs them
on seeded inputs, and commits inputs / parameters / outputs / gradients as small
fixtures under ``tests/golden/``; ``tests/test_oracle_golden.py`` holds the oracle to
those fixtures, and ``tests/test_oracle_vs_reference.py`` re-checks it live whenever
``/root/reference`` is present.
"""
