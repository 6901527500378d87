"""Host-side boundary checks that need no GPU: the C-ABI library loads and exports every
symbol include/pwgb.h declares, the module mirror reproduces the reference's state-dict
layout (names + shapes recorded from the real reference in the golden fixtures), and the
product path refuses CPU tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

from helpers import golden_weights, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__

    __graft_entry__.build()
    from parallelwavegan_b200 import capi

    return capi


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "pwgb.h")).read()
    declared = sorted(set(re.findall(r"\b(pwgb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found in include/pwgb.h"
    lib = ctypes.CDLL(built_lib.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"libpwgb.so lacks {missing}"
    assert sorted(built_lib.EXPORTED_SYMBOLS) == declared
    assert lib.pwgb_compiled_arch() == 100


def _mirror(kind, kwargs):
    from parallelwavegan_b200 import models

    cls = {"hifigan_generator": models.HiFiGANGenerator, "melgan_generator": models.MelGANGenerator,
           "pwg_generator": getattr(models, "ParallelWaveGANGenerator", None),
           "style_melgan_generator": models.StyleMelGANGenerator}[kind]
    if cls is None:
        pytest.skip(f"{kind} not built yet")
    import json

    return cls(**json.loads(json.dumps(kwargs)))


@pytest.mark.parametrize("name", ["hifigan_small", "hifigan_v1", "hifigan_odd", "hifigan_causal", "mb_melgan_v2", "melgan_small", "melgan_causal", "pwg_v1", "pwg_small", "style_melgan_v1", "style_melgan_small"])
def test_state_dict_layout_matches_reference(name, built_lib):
    meta, _ = load_golden(name)
    m = _mirror(meta["kind"], meta["kwargs"])
    ours = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    assert ours == [(k, list(s)) for k, s in meta["spec"]]
    # reference-layout checkpoints load strictly, with and without weight norm
    m.load_state_dict(golden_weights(meta), strict=True)
    m.remove_weight_norm()
    assert not any(k.endswith("weight_g") for k in m.state_dict())


def test_cpu_tensors_are_refused(built_lib):
    from parallelwavegan_b200 import models
    from parallelwavegan_b200.capi import PwgbError

    m = models.HiFiGANGenerator(channels=32)
    with pytest.raises(PwgbError, match="no CPU fallback"):
        m(torch.randn(1, 80, 8))
