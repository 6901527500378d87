"""Secondary inference measurements (not the headline bench): PWG v1, MB-MelGAN v2 + PQMF, HiFi-GAN batch 1,
MR-STFT loss at C3 size.  CUDA events, 3 warm-ups, L2 flushed between iterations."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from oracle import synth
from parallelwavegan_b200 import layers, losses, models, ops

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)


def timeit(fn, iters=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def load(m, seed, gain):
    m.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed, gain))
    m.remove_weight_norm()
    return m.eval().to(dev)


out = {}
with torch.no_grad():
    # PWG v1 generator (C1 shape and a batch-16 variant)
    pwg = load(models.ParallelWaveGANGenerator(), 31, 1.0)
    for B in (1, 16):
        c = torch.randn(B, 80, 104, device=dev)
        z = torch.randn(B, 1, 25600, device=dev)
        ms = timeit(lambda: pwg(z, c))
        out[f"pwg_v1_generator_B{B}x25600"] = {"ms": ms, "samples_per_s": B * 25600 / ms * 1e3, "x_realtime_24k": B * 25600 / ms * 1e3 / 24000}
    # MB-MelGAN v2 + PQMF (C4: 32 x 80 x 400 -> 32 x 1 x 120000)
    mb = load(models.MelGANGenerator(in_channels=80, out_channels=4, kernel_size=7, channels=384, upsample_scales=[5, 5, 3], stack_kernel_size=3, stacks=4), 21, 0.8)
    pq = layers.PQMF(4).to(dev)
    c = torch.randn(32, 80, 400, device=dev)
    ms = timeit(lambda: pq.synthesis(mb(c)))
    out["mb_melgan_v2_pqmf_B32x400"] = {"ms": ms, "samples_per_s": 32 * 120000 / ms * 1e3, "x_realtime_24k": 32 * 120000 / ms * 1e3 / 24000}
    # HiFi-GAN v1 batch 1 (RTF target of the north star), eager launches
    hf, _ = bench.synth_weights()
    hf = hf.to(dev)
    c1 = torch.randn(1, 80, 400, device=dev)
    ms = timeit(lambda: hf(c1), iters=10)
    out["hifigan_v1_B1x400"] = {"ms": ms, "samples_per_s": 102400 / ms * 1e3, "rtf": ms * 1e-3 / (102400 / 22050), "x_realtime": 102400 / ms * 1e3 / 22050}
    # MR-STFT loss forward at C3 size
    mr = losses.MultiResolutionSTFTLoss().to(dev)
    x = torch.rand(64, 25600, device=dev) - 0.5
    y = torch.rand(64, 25600, device=dev) - 0.5
    ms = timeit(lambda: mr(x, y))
    out["mr_stft_loss_64x25600"] = {"ms": ms, "alg_GBps": 8.0 * 64 * 25600 / ms / 1e6}
print(json.dumps(out, indent=1))
