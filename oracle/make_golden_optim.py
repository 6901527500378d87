#!/usr/bin/env python
"""Golden vectors for the optimizer step: runs the REAL reference RAdam (parallel_wavegan/optimizers/radam.py)
and torch.optim.Adam for 12 steps on seeded parameters / gradients (the RAdam rectification switches on at
step 6 for beta2 = 0.999) and stores the parameters after steps 1, 5, 6 and 12.  Build container only."""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import GOLD, import_reference  # noqa: E402

import torch  # noqa: E402

from oracle import synth  # noqa: E402

SHAPES = [(64, 32, 3), (64,), (7, 5), (1,), (130000,)]
STEPS = 12
KEEP = (1, 5, 6, 12)


def grads_for(step):
    return [synth.randn(s, 9000 + 10 * step + i, 0.5) for i, s in enumerate(SHAPES)]


def main():
    import_reference()
    from parallel_wavegan.optimizers import RAdam

    out = {}
    for name, mk in (("radam", lambda ps: RAdam(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0)),
                     ("radam_wd", lambda ps: RAdam(ps, lr=1e-2, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01)),
                     ("adam", lambda ps: torch.optim.Adam(ps, lr=2e-3, betas=(0.5, 0.9), eps=1e-8)),
                     ("adam_clip", lambda ps: torch.optim.Adam(ps, lr=2e-3, betas=(0.5, 0.9), eps=1e-8))):
        ps = [torch.nn.Parameter(synth.randn(s, 8000 + i)) for i, s in enumerate(SHAPES)]
        opt = mk(ps)
        for t in range(1, STEPS + 1):
            for p, g in zip(ps, grads_for(t)):
                p.grad = g.clone()
            if name.endswith("clip"):
                torch.nn.utils.clip_grad_norm_(ps, 3.0)
            opt.step()
            if t in KEEP:
                for i, p in enumerate(ps):
                    a = p.detach().numpy().copy()
                    out[f"{name}_t{t}_p{i}"] = a.reshape(-1)[:1024].copy() if a.size > 1024 else a
                    out[f"{name}_t{t}_s{i}"] = np.array([float(p.detach().double().sum()), float(p.detach().double().norm())])
    np.savez_compressed(os.path.join(GOLD, "optim.npz"), **out)
    print("wrote optim.npz", len(out), "arrays")
    make_data_golden()


def data_items():
    """Seeded synthetic corpus: (audio, mel) pairs of ragged lengths, one too short to be kept, one needing edge padding."""
    items = []
    for k, frames in enumerate((40, 25, 12, 33, 60)):
        x = synth.randn((frames * 64,), 7000 + k).numpy()
        if k == 3:
            x = x[:-17]  # _adjust_length edge-pads this one
        items.append((x, synth.randn((frames, 20), 7100 + k).numpy()))
    return items


def make_data_golden():
    """Collater golden: the REAL reference Collater (bin/train.py:646-925) on the seeded corpus with np.random.seed(11)."""
    # bin/train.py touches a few names of absent logging / plotting packages at import time: give the stub modules those names
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["matplotlib"].use = lambda *a, **k: None
    from parallel_wavegan.bin.train import Collater

    out = {}
    items = data_items()
    np.random.seed(11)
    (c,), y = Collater(batch_max_steps=1100, hop_size=64, aux_context_window=2, use_noise_input=False)(items)
    out["mel2wav_c"], out["mel2wav_y"] = c.numpy(), y.numpy()
    np.random.seed(12)
    (z, c2), y2 = Collater(batch_max_steps=512, hop_size=64, aux_context_window=0, use_noise_input=True)(items)
    out["noise_c"], out["noise_y"], out["noise_z_shape"] = c2.numpy(), y2.numpy(), np.array(z.shape)
    np.random.seed(13)
    (_l, _g), y3 = Collater(batch_max_steps=1500, hop_size=None, aux_context_window=0, use_aux_input=False)([x for x, _ in items])
    out["audio_y"] = y3.numpy()
    np.savez_compressed(os.path.join(GOLD, "data.npz"), **out)
    print("wrote data.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
