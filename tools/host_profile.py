"""cProfile of the host side of one HiFi-GAN (C5) training step: where the Python / ctypes time between launches goes."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from parallelwavegan_b200 import losses, models
from parallelwavegan_b200 import synth_weights as synth
from parallelwavegan_b200.optimizers import FusedAdam
from parallelwavegan_b200.train_step import GanTrainStep

dev = torch.device("cuda:0")
g = models.HiFiGANGenerator(**bench.CFG)
g.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in g.state_dict().items()], 1234, 1.15))
d = models.HiFiGANMultiScaleMultiPeriodDiscriminator()
d.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in d.state_dict().items()], 4321, 1.4))
g, d = g.to(dev).train(), d.to(dev).train()
crit = {"mel": losses.MelSpectrogramLoss(fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80, fmin=0, fmax=11025, log_base=None).to(dev),
        "gen_adv": losses.GeneratorAdversarialLoss(), "dis_adv": losses.DiscriminatorAdversarialLoss(), "feat_match": losses.FeatureMatchLoss()}
step = GanTrainStep(g, d, crit, FusedAdam(g.parameters(), lr=2e-4, betas=(0.5, 0.9)), FusedAdam(d.parameters(), lr=2e-4, betas=(0.5, 0.9)), steps=1)
c = torch.randn(16, 80, 32, device=dev)
y = torch.rand(16, 1, 8192, device=dev) - 0.5
for _ in range(3):
    step(c, y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step(c, y)
e1.record()
torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1) / 3)
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step(c, y)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
