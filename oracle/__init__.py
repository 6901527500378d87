"""CPU oracle for the vocoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``parallelwavegan_b200/`` may import
this package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and there only as the
checker / the CPU baseline -- never as the thing measured as "ours" or shipped.

Parity status: PINNED.  Every function in :mod:`oracle.ref_ops` is checked by
``tests/test_oracle_golden.py`` against golden vectors produced by importing the
real reference (``/root/reference/parallel_wavegan``) in the build container
with ``oracle/make_golden.py`` (committed; fixtures in ``tests/golden/``).
"""
