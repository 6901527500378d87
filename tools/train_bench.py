"""HiFi-GAN v1 G + MSD/MPD train step (BASELINE config C5: batch 16 x 8192 samples per GPU), steps/s.
Run single process or under torchrun (DDP over NCCL)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from oracle import synth
from parallelwavegan_b200 import losses, models, ops
from parallelwavegan_b200.train_step import GanTrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--profile", action="store_true")
args = ap.parse_args()
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
g = models.HiFiGANGenerator(**bench.CFG)
g.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in g.state_dict().items()], 1234, 1.15))
d = models.HiFiGANMultiScaleMultiPeriodDiscriminator()
d.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in d.state_dict().items()], 4321, 1.4))
g, d = g.to(dev).train(), d.to(dev).train()
if world > 1:
    g = torch.nn.parallel.DistributedDataParallel(g, device_ids=[local])
    d = torch.nn.parallel.DistributedDataParallel(d, device_ids=[local])
crit = {
    "mel": losses.MelSpectrogramLoss(fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80, fmin=0, fmax=11025, log_base=None).to(dev),
    "gen_adv": losses.GeneratorAdversarialLoss(), "dis_adv": losses.DiscriminatorAdversarialLoss(), "feat_match": losses.FeatureMatchLoss(),
}
opt_g = torch.optim.Adam(g.parameters(), lr=2e-4, betas=(0.5, 0.9))
opt_d = torch.optim.Adam(d.parameters(), lr=2e-4, betas=(0.5, 0.9))
step = GanTrainStep(g, d, crit, opt_g, opt_d)
gen = torch.Generator().manual_seed(rank)
c = torch.randn(args.batch, 80, 32, generator=gen).to(dev)
y = (torch.rand(args.batch, 1, 8192, generator=gen) - 0.5).to(dev)
for _ in range(args.warmup):
    st = step(c, y)
torch.cuda.synchronize()
if args.profile:
    ops.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    st = step(c, y)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "HiFi-GAN v1 G + MSD/MPD train step (mel + adv + feat-match, Adam), per-GPU batch %d x 8192" % args.batch,
                      "n_gpus": world, "ms_per_step": float(t[0]), "steps_per_sec": 1e3 / float(t[0]),
                      "losses": {k: float(v) for k, v in st.items()}}))
    if args.profile:
        agg = {}
        for name, fl, by, a, b, desc in ops.PROFILE:
            v = agg.setdefault(name, [0.0, 0])
            v[0] += a.elapsed_time(b)
            v[1] += 1
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:20s} {v[0] / args.steps:9.2f} ms/step {v[1] // args.steps:5d} launches/step")
if world > 1:
    dist.destroy_process_group()
