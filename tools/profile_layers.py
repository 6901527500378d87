"""Per-launch table of the HiFi-GAN v1 C2 forward (CUDA events on the launching stream)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from parallelwavegan_b200 import ops

model, _ = bench.synth_weights()
dev = torch.device("cuda:0")
model = model.to(dev)
mel = torch.randn(bench.BATCH, 80, bench.FRAMES, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(mel)
    torch.cuda.synchronize()
    ops.PROFILE = []
    model(mel)
    torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
tot = 0.0
print(f"{'kernel':18s} {'shape':46s} {'ms':>8s} {'TFLOP/s':>8s} {'algGB/s':>8s}")
for name, fl, by, a, b, desc in prof:
    ms = a.elapsed_time(b)
    tot += ms
    print(f"{name:18s} {desc:46s} {ms:8.3f} {fl / ms / 1e9:8.1f} {by / ms / 1e6:8.0f}")
print("total ms", tot)
