"""Per-role timeline of CTA 0 of the fused WaveNet kernel (pwgb_debug_set(2, 64) + pwgb_debug_get(2)).
usage: wn_trace.py [dilation,T,B] [extra variant bits]   -- prints SM-clock stamps relative to the first one."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from parallelwavegan_b200 import capi, layers, ops
from parallelwavegan_b200 import synth_weights as synth

dev = torch.device("cuda:0")
d, T, B = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "1,25600,16").split(",")]
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 0
blk = layers.WaveNetResidualBlock(dilation=d)
blk.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in blk.state_dict().items()], 40 + d, 1.0))
blk = blk.to(dev).eval()
x = torch.randn(B, 64, T, device=dev)
c = torch.zeros(B, 96, T, device=dev)
c[:, :80].normal_()
skips = torch.zeros(B, 64, T, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
st = ops.WnStack(B, T, 64, 128, 64, 80, 3, 512, dev)
st.pack_c(c)
st.pack_x(x)
with torch.no_grad():
    packed, bso = ops.wavenet_packed_weights(layers.effective_weight(blk.conv), layers.effective_weight(blk.conv1x1_aux),
                                             layers.effective_weight(blk.conv1x1_skip), layers.effective_weight(blk.conv1x1_out),
                                             blk.conv1x1_skip.bias, blk.conv1x1_out.bias, 80)
    for _ in range(3):
        st.layer(packed, blk.conv.bias, bso, d, skips)
    capi.lib().pwgb_debug_set(2, 64 | extra)
    flush.zero_()
    st.layer(packed, blk.conv.bias, bso, d, skips)
    torch.cuda.synchronize()
    capi.lib().pwgb_debug_set(2, 0)
buf = np.zeros((8, 32, 4), dtype=np.int64)
n = capi.lib().pwgb_debug_get(2, C.c_void_p(buf.ctypes.data), buf.nbytes)
assert n == buf.nbytes, n
big = buf > (1 << 32)  # clock64 stamps; the small entries are accumulated wait cycles
t0 = buf[big].min()
rel = np.where(big, buf - t0, np.where(buf > 0, buf, -1))
names = ["epi-skip", "epi-x", "gate", "so-issuer", "conv0", "conv1", "loader", "-"]
stamps = {0: ["start", "so_full", "done", "store_wait"], 1: ["start", "so_full", "done", "store_wait"], 2: ["start", "g_full", "z_empty", "done"],
          3: ["start", "z_full", "issued"], 4: ["acc_empty", "conv_issued", "full_wait"], 5: ["acc_empty", "conv_issued", "full_wait"], 6: ["start", "conv_loaded", "empty_wait"]}
print(f"variant bits {64 | extra}; d{d} T{T} B{B}; clocks relative to the first stamp")
for r in (6, 4, 5, 2, 3, 0, 1):
    print(f"== {names[r]}: " + ", ".join(stamps[r]))
    for t in range(24):
        row = rel[r, t]
        if (row[:2] >= 0).any():
            print(f"  tile {t:2d}: " + "  ".join(f"{int(v):8d}" for v in row[: len(stamps[r])]))
# per-tile period (steady state) from the epilogue done stamps
done = np.sort(np.concatenate([rel[0, :, 2], rel[1, :, 2]]))
done = done[done > 0][::2]
if len(done) > 6:
    print("steady-state period per tile (clocks):", float(np.diff(done[2:]).mean()))
