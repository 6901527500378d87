"""Pins the optimizer restatement (oracle/ref_optim.py) to golden vectors produced by the REAL reference
RAdam (parallel_wavegan/optimizers/radam.py) and by torch.optim.Adam (oracle/make_golden_optim.py)."""
import os

import numpy as np
import torch

from helpers import GOLD
from oracle import ref_optim, synth
from oracle.make_golden_optim import KEEP, SHAPES, STEPS, grads_for


def _run(kind, lr, betas, eps, wd=0.0, clip=None):
    ps = [synth.randn(s, 8000 + i) for i, s in enumerate(SHAPES)]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    snaps = {}
    for t in range(1, STEPS + 1):
        gs = grads_for(t)
        if clip:
            _, c = ref_optim.clip_coef(gs, clip)
            gs = [g * c for g in gs]
        for p, g, m, v in zip(ps, gs, ms, vs):
            (ref_optim.radam_step if kind == "radam" else ref_optim.adam_step)(p, g, m, v, t, lr, betas, eps, wd)
        if t in KEEP:
            snaps[t] = [p.clone() for p in ps]
    return snaps


def test_optimizer_restatement_matches_reference():
    g = np.load(os.path.join(GOLD, "optim.npz"))
    runs = {"radam": _run("radam", 1e-2, (0.9, 0.999), 1e-6), "radam_wd": _run("radam", 1e-2, (0.9, 0.999), 1e-6, 0.01),
            "adam": _run("adam", 2e-3, (0.5, 0.9), 1e-8), "adam_clip": _run("adam", 2e-3, (0.5, 0.9), 1e-8, clip=3.0)}
    for name, snaps in runs.items():
        for t, ps in snaps.items():
            for i, p in enumerate(ps):
                ref = torch.from_numpy(g[f"{name}_t{t}_p{i}"])
                got = p.reshape(-1)[: ref.numel()].reshape(ref.shape)
                assert float((got - ref).abs().max()) <= 2e-6 * max(1.0, float(ref.abs().max())), (name, t, i)
                s = g[f"{name}_t{t}_s{i}"]
                assert abs(float(p.double().sum()) - s[0]) <= 1e-5 * max(1.0, s[1] * p.numel() ** 0.5)
                assert abs(float(p.double().norm()) - s[1]) <= 1e-5 * max(1.0, s[1])
