"""PWG v1 generator forward (30 fused layers on the packed stream) timing: B x 25600, L2 flushed, CUDA events;
prints samples/s and the fraction of the HBM roofline of the residual stack (30 x 1344 B per sample)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from parallelwavegan_b200 import models, ops
from parallelwavegan_b200 import synth_weights as synth

dev = torch.device("cuda:0")
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
m = models.ParallelWaveGANGenerator()
m.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], 31, 1.0))
m.remove_weight_norm()
m = m.eval().to(dev)
out = {}
with torch.no_grad():
    for B in [int(a) for a in (sys.argv[1:] or ["1", "16", "64"])]:
        T = 25600
        z = torch.randn(B, 1, T, device=dev)
        c = torch.randn(B, 80, T // 256 + 4, device=dev)
        for _ in range(3):
            m(z, c)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m(z, c)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        ops.PROFILE = []
        m(z, c)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for name, fl, by, a, b, _d in prof:
            e = agg.setdefault(name, [0.0, 0])
            e[0] += a.elapsed_time(b)
            e[1] += 1
        rate = B * T / (ms * 1e-3)
        fused_ms = agg.get("wavenet_fused_tc", [0.0, 0])[0]
        out[f"B{B}"] = {"ms": ms, "samples_per_s": rate, "stack_alg_GBps": rate * 30 * 1344 / 1e9, "stack_frac_of_hbm_whole_forward": rate * 30 * 1344 / 1e9 / HBM,
                       "fused_layers_ms": fused_ms, "fused_layers_frac_of_hbm": (B * T * (30 * 1344 - 256 - 256) / 1e9) / (fused_ms * 1e-3) / HBM if fused_ms else None,
                       "classes": {k: [round(v[0], 3), v[1]] for k, v in agg.items()}}
print(json.dumps(out, indent=1))
