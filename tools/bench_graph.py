"""Batch-1 HiFi-GAN v1 latency: eager launches vs CUDA-graph replay (decode driver)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import json

import torch

import bench
from parallelwavegan_b200.decode import GraphedGenerator

dev = torch.device("cuda:0")
m, _ = bench.synth_weights()
m = m.to(dev)
c = torch.randn(1, 80, 400, device=dev)
gm = GraphedGenerator(m)
res = {}
with torch.no_grad():
    for name, fn in (("eager", lambda: m(c)), ("cuda_graph", lambda: gm(c))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[name] = {"ms": ms, "rtf": ms * 1e-3 / (102400 / 22050), "x_realtime": 102400 / 22050 / (ms * 1e-3)}
    assert torch.equal(m(c), gm(c))
print(json.dumps({"hifigan_v1_batch1_400frames": res}))
