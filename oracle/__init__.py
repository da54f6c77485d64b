"""CPU oracle for the two DCTdomain hot paths.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or the timed CPU baseline), never
as a fallback for the CUDA path.  ``dctdomain_b200`` never imports this package.

Parity status:
  * fingerprint path  - PINNED: ``fingerprint_oracle`` is checked against outputs of
    the unmodified reference ``src/fingerprint.py`` / ``src/embedding.py`` generated in
    the build container by ``tests/golden/make_golden.py`` (fixtures in ``tests/golden``).
  * search path       - PINNED on ordering / padding by the reference's own
    ``test/test/example-search.txt`` (400 lines) and ``bench/G6PD/G6PD-dctsim.txt``;
    the k-th-boundary tie rule (k smallest by (dist, id)) follows faiss 1.7.4
    ``IndexFlat`` + ``METRIC_L1`` heap semantics, which no reference fixture exercises
    (N=43 < k=50 there): that one rule is "parity unpinned".
"""
