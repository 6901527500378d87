"""Run single WaveNet residual layers (PWG v1 sizes) for ncu captures / quick timing:
args = list of 'dilation,T,B'.  Prints ms, sample-layers/s and the fraction of the HBM roofline
(1344 algorithmic bytes per sample-layer) against MEASURED_PEAKS.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from parallelwavegan_b200 import layers, ops
from parallelwavegan_b200 import synth_weights as synth

dev = torch.device("cuda:0")
if os.environ.get("PWGB_WN_VARIANT"):
    from parallelwavegan_b200 import capi

    capi.lib().pwgb_debug_set(2, int(os.environ["PWGB_WN_VARIANT"]))  # timing experiments: outputs are not valid
    print("WN variant", os.environ["PWGB_WN_VARIANT"])
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM = 6650.0
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
for spec in sys.argv[1:]:
    d, T, B = [int(v) for v in spec.split(",")]
    blk = layers.WaveNetResidualBlock(dilation=d)
    blk.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in blk.state_dict().items()], 40 + d, 1.0))
    blk = blk.to(dev).eval()
    x = torch.randn(B, 64, T, device=dev)
    c = torch.zeros(B, 96, T, device=dev)
    c[:, :80].normal_()
    skips = torch.zeros(B, 64, T, device=dev)
    with torch.no_grad():
        for _ in range(3):
            blk(x, c, skips)
        torch.cuda.synchronize()
        evs = []
        for _ in range(5):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            blk(x, c, skips)
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[2]
    rate = B * T / (ms * 1e-3)
    if os.environ.get("WN_FUSED", "1") != "0" and ops.WnStack.supported(B, T, 64, 128, 64, 80, 3, 512):
        st = ops.WnStack(B, T, 64, 128, 64, 80, 3, 512, dev)
        st.pack_c(c)
        st.pack_x(x)
        with torch.no_grad():
            packed, bso = ops.wavenet_packed_weights(layers.effective_weight(blk.conv), layers.effective_weight(blk.conv1x1_aux),
                                                     layers.effective_weight(blk.conv1x1_skip), layers.effective_weight(blk.conv1x1_out),
                                                     blk.conv1x1_skip.bias, blk.conv1x1_out.bias, 80)
            for _ in range(3):
                st.layer(packed, blk.conv.bias, bso, d, skips)
            torch.cuda.synchronize()
            evs = []
            for _ in range(5):
                flush.zero_()
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                st.layer(packed, blk.conv.bias, bso, d, skips)
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
        msf = sorted(a.elapsed_time(b) for a, b in evs)[2]
        rf = B * T / (msf * 1e-3)
        print(f"FUSED   layer d{d} T{T} B{B}: {msf:.3f} ms  {rf / 1e9:.2f} G sample-layers/s  {rf * 1344 / 1e9:.0f} GB/s(alg) = "
              f"{rf * 1344 / 1e9 / HBM:.3f} of measured HBM {HBM:.0f} GB/s")
    print(f"2-launch layer d{d} T{T} B{B}: {ms:.3f} ms  {rate / 1e9:.2f} G sample-layers/s  {rate * 1344 / 1e9:.0f} GB/s(alg) = "
          f"{rate * 1344 / 1e9 / HBM:.3f} of measured HBM {HBM:.0f} GB/s")
