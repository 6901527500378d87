"""Deterministic synthetic checkpoints: random-init weights in the reference's state-dict layout.

There is no network here, so neither the reference's pretrained checkpoints nor its datasets are
reachable, and its constructor inits (N(0, 0.01) for HiFi-GAN, hifigan.py:202-205) make activations
vanish.  ``bench.py`` and the parity tests therefore draw weights with :func:`synth_state_dict` from a
seeded CPU generator (``oracle/synth.py`` re-exports these functions, and ``oracle/make_golden.py`` loads
exactly these tensors into the *reference* modules to produce the golden outputs).  Pure data
generation: no model arithmetic lives here.
"""

import math

import torch


def synth_state_dict(spec, seed, gain=1.0):
    """spec: list of (name, shape) in state_dict order.  Rules by key suffix:
    ``weight_v``/``weight``/``weight_orig``: N(0,1) * gain / sqrt(fan_in);
    ``weight_g``: ||v|| * (1 + 0.05 N(0,1)) (shape as given);
    ``bias``: 0.05 N(0,1); ``weight_u``: normalised N(0,1); ``mean``: 0.1 N;
    ``scale``: 1 + 0.1 U; other buffers (filters) must be supplied by the caller."""
    g = torch.Generator().manual_seed(int(seed))
    sd = {}
    for name, shape in spec:
        shape = tuple(int(s) for s in shape)
        leaf = name.rsplit(".", 1)[-1]
        if leaf in ("weight", "weight_v", "weight_orig"):
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            sd[name] = torch.randn(shape, generator=g) * (gain / math.sqrt(max(fan_in, 1)))
        elif leaf == "weight_g":
            v = sd.get(name[: -len("_g")] + "_v")
            noise = 1.0 + 0.05 * torch.randn(shape, generator=g)
            if v is not None:
                dims = tuple(range(1, v.dim()))
                sd[name] = v.pow(2).sum(dim=dims, keepdim=True).sqrt().reshape(shape) * noise
            else:
                sd[name] = noise
        elif leaf == "bias":
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif leaf == "weight_u":
            u = torch.randn(shape, generator=g)
            sd[name] = u / u.norm()
        elif leaf == "mean":
            sd[name] = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "scale":
            sd[name] = 1.0 + 0.1 * torch.rand(shape, generator=g)
        else:
            raise KeyError(f"no synthesis rule for {name}")
    # weight_g must be generated after its weight_v: state_dict order is g then v
    for name, shape in spec:
        if name.endswith(".weight_g"):
            v = sd[name[: -len("_g")] + "_v"]
            dims = tuple(range(1, v.dim()))
            g2 = torch.Generator().manual_seed(int(seed) * 7919 + (sum(name.encode()) % 65521))
            noise = 1.0 + 0.05 * torch.randn(tuple(shape), generator=g2)
            sd[name] = v.pow(2).sum(dim=dims, keepdim=True).sqrt().reshape(tuple(shape)) * noise
    return sd


def checksum(sd):
    """Order-dependent float64 checksum of a state dict (detects RNG drift)."""
    tot = 0.0
    for i, (k, v) in enumerate(sd.items()):
        tot += float(v.double().abs().sum()) * (1.0 + (i % 17) * 1e-3)
    return tot


def randn(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(int(seed))
    return torch.randn(tuple(shape), generator=g) * scale
