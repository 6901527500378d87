"""CPU restatement of the optimizer side of Trainer._train_step (TEST INFRASTRUCTURE -- only tests/,
smoke() and bench.py's CPU legs may import oracle/).

* ``radam_step``  -- one step of the reference's RAdam (parallel_wavegan/optimizers/radam.py:27-99) on plain
  tensors; pinned against the real reference by tests/golden/optim.npz (oracle/make_golden_optim.py).
* ``adam_step``   -- torch.optim.Adam (amsgrad=False) semantics, the optimizer the HiFi-GAN recipes name
  (egs/ljspeech/voc1/conf/hifigan.v1.yaml:136-163); pinned against torch.optim.Adam on CPU in the tests.
* ``clip_coef``   -- torch.nn.utils.clip_grad_norm_ (bin/train.py:289-293): global L2 norm, max_norm/(norm+1e-6) <= 1.
"""
import math

import torch


def clip_coef(grads, max_norm):
    norm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    return norm, torch.clamp(max_norm / (norm + 1e-6), max=1.0)


def radam_step(p, g, m, v, t, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """In-place on (p, m, v); t = step count AFTER this step (radam.py:58)."""
    b1, b2 = betas
    v.mul_(b2).addcmul_(g, g, value=1 - b2)  # radam.py:55
    m.mul_(b1).add_(g, alpha=1 - b1)  # radam.py:56
    beta2_t = b2**t
    n_sma_max = 2 / (1 - b2) - 1
    n_sma = n_sma_max - 2 * t * beta2_t / (1 - beta2_t)  # radam.py:65-66
    if n_sma >= 5:
        step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max / (n_sma_max - 2)) / (1 - b1**t)
    else:
        step_size = 1.0 / (1 - b1**t)
    if weight_decay != 0:
        p.add_(p, alpha=-weight_decay * lr)  # radam.py:86
    if n_sma >= 5:
        p.addcdiv_(m, v.sqrt().add_(eps), value=-step_size * lr)  # radam.py:90-91
    else:
        p.add_(m, alpha=-step_size * lr)  # radam.py:93
    return p


def adam_step(p, g, m, v, t, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    b1, b2 = betas
    if weight_decay != 0:
        g = g.add(p, alpha=weight_decay)
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    denom = (v.sqrt() / math.sqrt(1 - b2**t)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / (1 - b1**t))
    return p
