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


if __name__ == "__main__":
    main()
