"""TEST INFRASTRUCTURE ONLY — CPU restatement of the CTGCN hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the
thing timed *as the CPU baseline*), never as the code path that ships.

Parity status: the reference's own tests hold no golden vector for this path
 → the reference itself is "parity unpinned".  This oracle
is pinned instead against outputs of the unmodified reference modules
(``/root/reference/layers.py`` / ``models.py``) executed in the build
container: ``oracle/make_golden.py`` generates ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` re-checks both restatements against them.
"""
