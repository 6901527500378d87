"""Parallel WaveGAN v1 G + D train step (BASELINE config C3: MR-STFT + adversarial loss, RAdam), steps/s."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import synth
from parallelwavegan_b200 import losses, models, ops

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
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
g = models.ParallelWaveGANGenerator()
g.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in g.state_dict().items()], 31, 1.0))
d = models.ParallelWaveGANDiscriminator()
d.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in d.state_dict().items()], 64, 1.4))
g, d = g.to(dev).train(), d.to(dev).train()
if world > 1:
    g = torch.nn.parallel.DistributedDataParallel(g, device_ids=[local], find_unused_parameters=True)
    d = torch.nn.parallel.DistributedDataParallel(d, device_ids=[local])
mr = losses.MultiResolutionSTFTLoss().to(dev)
gen_adv, dis_adv = losses.GeneratorAdversarialLoss(), losses.DiscriminatorAdversarialLoss()
opt_g = torch.optim.RAdam(g.parameters(), lr=1e-4, eps=1e-6)
opt_d = torch.optim.RAdam(d.parameters(), lr=5e-5, eps=1e-6)
B, T = args.batch, 25600
gen = torch.Generator().manual_seed(rank)
c = torch.randn(B, 80, T // 256 + 4, generator=gen).to(dev)
y = ((torch.rand(B, 1, T, generator=gen) - 0.5)).to(dev)


def dparams():
    return (d.module if hasattr(d, "module") else d).parameters()


def step():
    z = torch.randn(B, 1, T, device=dev)
    # generator phase (train.py:200-295): MR-STFT aux loss + lambda_adv * adversarial
    y_ = g(z, c)
    sc, mag = mr(y_.squeeze(1), y.squeeze(1))
    for p in dparams():
        p.requires_grad_(False)
    adv = gen_adv(d(y_))
    gen_loss = sc + mag + 4.0 * adv
    opt_g.zero_grad(set_to_none=True)
    gen_loss.backward()
    torch.nn.utils.clip_grad_norm_(g.parameters(), 10.0)
    opt_g.step()
    for p in dparams():
        p.requires_grad_(True)
    # discriminator phase (train.py:300-335)
    with torch.no_grad():
        y_ = g(z, c)
    real, fake = dis_adv(d(y_.detach()), d(y))
    dis_loss = real + fake
    opt_d.zero_grad(set_to_none=True)
    dis_loss.backward()
    torch.nn.utils.clip_grad_norm_(d.parameters(), 1.0)
    opt_d.step()
    return gen_loss.detach(), dis_loss.detach()


for _ in range(args.warmup):
    st = step()
torch.cuda.synchronize()
if args.profile:
    ops.PROFILE = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    st = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": f"ParallelWaveGAN v1 G + D train step (MR-STFT + adv, RAdam), per-GPU batch {B} x {T}", "n_gpus": world,
                      "ms_per_step": float(t[0]), "steps_per_sec": 1e3 / float(t[0]), "gen_loss": float(st[0]), "dis_loss": float(st[1]),
                      "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}))
    if args.profile:
        agg = {}
        for name, fl, by, a, b, desc in ops.PROFILE:
            v = agg.setdefault(name, [0.0, 0])
            v[0] += a.elapsed_time(b)
            v[1] += 1
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            print(f"  {k:20s} {v[0] / args.steps:9.2f} ms/step {v[1] // args.steps:5d} launches/step")
        agg2 = {}
        for name, fl, by, a, b, desc in ops.PROFILE:
            v = agg2.setdefault(name + " " + desc, [0.0, 0])
            v[0] += a.elapsed_time(b)
            v[1] += 1
        for k, v in sorted(agg2.items(), key=lambda kv: -kv[1][0])[:12]:
            print(f"    {v[0] / args.steps:9.2f} ms/step x{v[1] // args.steps:3d}  {k}")
if world > 1:
    dist.destroy_process_group()
