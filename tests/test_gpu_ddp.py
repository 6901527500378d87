"""Data-parallel training numerics: two ranks, each with half of a batch, wrapped in
``torch.nn.parallel.DistributedDataParallel`` -- the averaged gradients must equal the gradients of ONE process on the
whole batch (the reference's apex DDP has exactly this mean-over-ranks semantics, train.py:1494-1503).  Both ranks share
cuda:0 here (gloo moves the gradient buckets; the round-end driver exercises NCCL over NVLink in the scaling runs)."""
import os
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu

KW = dict(in_channels=80, out_channels=1, channels=64, kernel_size=7, upsample_scales=[8, 8, 2, 2],
          upsample_kernel_sizes=[16, 16, 4, 4], resblock_kernel_sizes=[3, 7], resblock_dilations=[[1, 3, 5], [1, 3]])


DKW = dict(in_channels=1, out_channels=1, kernel_sizes=[5, 3], channels=8, downsample_scales=[3, 3, 1], max_downsample_channels=64, bias=True,
           nonlinear_activation="LeakyReLU", nonlinear_activation_params={"negative_slope": 0.1}, use_weight_norm=True, use_spectral_norm=False)


def _build(dev):
    from oracle import synth
    from parallelwavegan_b200 import losses, models

    g = models.HiFiGANGenerator(**KW)
    g.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in g.state_dict().items()], 7, 1.15))
    d = models.HiFiGANMultiPeriodDiscriminator(periods=[2, 3], discriminator_params=DKW)
    d.load_state_dict(synth.synth_state_dict([(k, tuple(v.shape)) for k, v in d.state_dict().items()], 9, 1.4))
    mel = losses.MelSpectrogramLoss(fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80, fmin=0,
                                    fmax=11025, log_base=None).to(dev)
    return g.to(dev).train(), d.to(dev).train(), mel


def _loss(g, d, mel, c, y):
    from parallelwavegan_b200 import losses

    y_ = g(c)
    p_ = d(y_)
    with torch.no_grad():
        p = d(y)
    return 45.0 * mel(y_, y) + losses.GeneratorAdversarialLoss()(p_) + 2.0 * losses.FeatureMatchLoss()(p_, p)


def _worker(rank, world, init_file, out_file):
    import torch.distributed as dist

    from oracle import synth

    import __graft_entry__

    __graft_entry__.build()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    g, d, mel = _build(dev)
    c = synth.randn((4, 80, 16), 21).to(dev)
    y = synth.randn((4, 1, 16 * 256), 22, 0.3).to(dev)
    gd = torch.nn.parallel.DistributedDataParallel(g)
    dd = torch.nn.parallel.DistributedDataParallel(d)
    lo, hi = rank * 2, rank * 2 + 2
    _loss(gd, dd, mel, c[lo:hi], y[lo:hi]).backward()
    if rank == 0:
        torch.save({"g": {k: p.grad.cpu() for k, p in g.named_parameters()}, "d": {k: p.grad.cpu() for k, p in d.named_parameters()}}, out_file)
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_gradients_equal_single_process_large_batch():
    assert torch.cuda.is_available()
    import torch.multiprocessing as mp

    from helpers import rel_l2
    from oracle import synth

    import __graft_entry__

    __graft_entry__.build()
    tmp = tempfile.mkdtemp()
    init_file, out_file = os.path.join(tmp, "rdzv"), os.path.join(tmp, "grads.pt")
    mp.spawn(_worker, args=(2, init_file, out_file), nprocs=2, join=True)
    ddp = torch.load(out_file)
    dev = torch.device("cuda:0")
    g, d, mel = _build(dev)
    c = synth.randn((4, 80, 16), 21).to(dev)
    y = synth.randn((4, 1, 16 * 256), 22, 0.3).to(dev)
    _loss(g, d, mel, c, y).backward()
    n = 0
    for name, mod in (("g", g), ("d", d)):
        for k, p in mod.named_parameters():
            # same kernels, same values: only the order of the batch reduction differs (two halves averaged vs one sum)
            assert rel_l2(ddp[name][k], p.grad.cpu()) < 1e-5, (name, k)
            n += 1
    assert n > 50
