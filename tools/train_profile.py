"""Top launches of one HiFi-GAN (default) or Parallel WaveGAN (`pwg [batch]`) train step (CUDA events per libpwgb launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections

import torch

import bench
from parallelwavegan_b200 import ops

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
# reuse the bench helper with instrumentation on
import parallelwavegan_b200.train_step  # noqa: F401

ops.PROFILE = None
res = None
orig = bench.measure_train_step


def run():
    from parallelwavegan_b200 import synth_weights as synth
    from parallelwavegan_b200 import losses, models
    from parallelwavegan_b200.train_step import GanTrainStep

    g = models.HiFiGANGenerator(**bench.CFG)
    g.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in g.state_dict().items()], 1234, 1.15))
    d = models.HiFiGANMultiScaleMultiPeriodDiscriminator()
    d.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in d.state_dict().items()], 4321, 1.4))
    g, d = g.to(dev).train(), d.to(dev).train()
    crit = {"mel": losses.MelSpectrogramLoss(fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80, fmin=0, fmax=11025, log_base=None).to(dev),
            "gen_adv": losses.GeneratorAdversarialLoss(), "dis_adv": losses.DiscriminatorAdversarialLoss(), "feat_match": losses.FeatureMatchLoss()}
    from parallelwavegan_b200.optimizers import FusedAdam
    step = GanTrainStep(g, d, crit, FusedAdam(g.parameters(), lr=2e-4, betas=(0.5, 0.9)), FusedAdam(d.parameters(), lr=2e-4, betas=(0.5, 0.9)), steps=1)
    c = torch.randn(16, 80, 32, device=dev)
    y = (torch.rand(16, 1, 8192, device=dev) - 0.5)
    for _ in range(2):
        step(c, y)
    torch.cuda.synchronize()
    if len(sys.argv) > 1 and sys.argv[1] == "torchprof":
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof_:
            step(c, y)
            torch.cuda.synchronize()
        print(prof_.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
        return
    ops.PROFILE = []
    step(c, y)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0])
    for name, fl, by, a, b, desc in prof:
        k = f"{name:18s} {desc}"
        agg[k][0] += a.elapsed_time(b)
        agg[k][1] += 1
        agg[k][2] += fl
    tot = sum(v[0] for v in agg.values())
    print("total instrumented ms", tot)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{v[0]:8.2f} ms x{v[1]:3d} {v[2] / max(v[0], 1e-9) / 1e9:7.1f} TF  {k}")


def run_pwg(batch):
    from parallelwavegan_b200 import synth_weights as synth
    from parallelwavegan_b200 import losses, models
    from parallelwavegan_b200.optimizers import RAdam
    from parallelwavegan_b200.train_step import GanTrainStep

    g = models.ParallelWaveGANGenerator()
    g.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in g.state_dict().items()], 31, 1.0))
    d = models.ParallelWaveGANDiscriminator()
    d.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in d.state_dict().items()], 64, 1.4))
    g, d = g.to(dev).train(), d.to(dev).train()
    crit = {"stft": losses.MultiResolutionSTFTLoss().to(dev), "gen_adv": losses.GeneratorAdversarialLoss(), "dis_adv": losses.DiscriminatorAdversarialLoss()}
    tstep = GanTrainStep(g, d, crit, RAdam(g.parameters(), lr=1e-4, eps=1e-6), RAdam(d.parameters(), lr=5e-5, eps=1e-6), lambda_aux=1.0,
                         lambda_adv=4.0, grad_norm_g=10.0, grad_norm_d=1.0, steps=1)
    T = 25600
    c = torch.randn(batch, 80, T // 256 + 4, device=dev)
    y = torch.rand(batch, 1, T, device=dev) - 0.5
    for _ in range(2):
        tstep((torch.randn(batch, 1, T, device=dev), c), y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tstep((torch.randn(batch, 1, T, device=dev), c), y)
    e1.record()
    torch.cuda.synchronize()
    print("uninstrumented step ms", e0.elapsed_time(e1))
    if len(sys.argv) > 3 and sys.argv[3] == "torchprof":
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof_:
            tstep((torch.randn(batch, 1, T, device=dev), c), y)
            torch.cuda.synchronize()
        print(prof_.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
        return
    ops.PROFILE = []
    tstep((torch.randn(batch, 1, T, device=dev), c), y)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
    for name, fl, by, a, b, desc in prof:
        k = f"{name:18s} {desc}"
        agg[k][0] += a.elapsed_time(b)
        agg[k][1] += 1
        agg[k][2] += fl
        agg[k][3] += by
    tot = sum(v[0] for v in agg.values())
    print("total instrumented ms", tot, "launches", len(prof))
    byname = collections.defaultdict(lambda: [0.0, 0])
    for name, fl, by, a, b, desc in prof:
        byname[name][0] += a.elapsed_time(b)
        byname[name][1] += 1
    for k, v in sorted(byname.items(), key=lambda kv: -kv[1][0]):
        print(f"   {v[0]:8.2f} ms x{v[1]:4d}  {k}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{v[0]:8.2f} ms x{v[1]:3d} {v[2] / max(v[0], 1e-9) / 1e9:7.1f} TF {v[3] / max(v[0], 1e-9) / 1e6:7.0f} GB/s  {k}")


if len(sys.argv) > 1 and sys.argv[1] == "pwg":
    run_pwg(int(sys.argv[2]) if len(sys.argv) > 2 else 64)
else:
    run()
