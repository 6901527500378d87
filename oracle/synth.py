"""Deterministic synthetic checkpoints (TEST INFRASTRUCTURE).

The generator itself lives in ``parallelwavegan_b200/synth_weights.py`` (``bench.py`` needs random-init
weights too and must not import ``oracle/``); this module re-exports it so that the fixtures, which store
``(spec, seed, gain)`` and a checksum, keep regenerating bit-identical tensors.
"""

from parallelwavegan_b200.synth_weights import checksum, randn, synth_state_dict  # noqa: F401
