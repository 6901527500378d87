"""GPU parity: discriminator towers and losses (through the C ABI) vs golden vectors from the
real reference and vs the CPU oracle.  Tolerance: rel <= 1e-3 (north_star)."""
import json

import pytest
import torch

from helpers import REL_TOL, check_fingerprint, golden_weights, load_golden, rel_l2
from oracle import ref_ops, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__

    __graft_entry__.build()
    return torch.device("cuda:0")


def _disc(name, dev):
    from parallelwavegan_b200 import models

    meta, g = load_golden(name)
    m = getattr(models, meta["cls"])(**json.loads(json.dumps(meta["kwargs"])))
    ours = [(k, list(v.shape)) for k, v in m.state_dict().items() if not k.endswith("_filter")]  # PQMF banks: fixed buffers
    assert ours == [(k, list(s)) for k, s in meta["spec"]], "state-dict layout differs from the reference"
    m.load_state_dict(golden_weights(meta), strict=not any(k.endswith("_filter") for k in m.state_dict()))
    m.train(meta["train_mode"])
    x = synth.randn(meta["x_shape"], meta["x_seed"], meta["x_scale"])
    return meta, g, m.to(dev), x


@pytest.mark.parametrize("name", ["hifigan_msmpd_v1", "hifigan_msmpd_v1_train", "melgan_msd", "pwg_disc", "style_melgan_disc"])
def test_discriminator_vs_reference(dev, name):
    import numpy as np

    meta, g, m, x = _disc(name, dev)
    if meta.get("np_seed") is not None:  # random-window discriminator: the reference's host RNG sequence (style_melgan.py:330)
        np.random.seed(meta["np_seed"])
    with torch.no_grad():
        outs = m(x.to(dev))
    check_fingerprint(outs, meta, g, REL_TOL)
    finals = [o[-1] if isinstance(o, (list, tuple)) else o for o in (outs if isinstance(outs, list) else [outs])]
    for i, f in enumerate(finals):
        assert rel_l2(f.cpu(), g[f"final{i}"]) < REL_TOL
    if meta["train_mode"]:  # spectral-norm power iteration state must advance exactly like the reference
        sd = m.state_dict()
        n = 0
        for k, v in sd.items():
            if k.endswith("weight_u"):
                assert rel_l2(v.cpu(), g["u__" + k.replace(".", "__")]) < 1e-4
                n += 1
        assert n == 8


def test_losses_vs_reference(dev):
    from parallelwavegan_b200 import losses

    _, g = load_golden("losses")
    x = synth.randn((3, 8192), 501, 0.3)
    y = synth.randn((3, 8192), 502, 0.3)
    y = 0.7 * y + 0.3 * x
    xd, yd = x.to(dev), y.to(dev)
    mr = losses.MultiResolutionSTFTLoss().to(dev)
    sc, mag = mr(xd, yd)
    assert sc.dim() == 0 and sc.is_cuda
    assert rel_l2(torch.stack([sc, mag]).cpu(), g["mr_default"]) < REL_TOL
    sc, mag = mr(xd.view(1, 3, -1), yd.view(1, 3, -1))
    assert rel_l2(torch.stack([sc, mag]).cpu(), g["mr_3d"]) < REL_TOL
    mr2 = losses.MultiResolutionSTFTLoss([64, 128, 256], [16, 32, 64], [64, 128, 256]).to(dev)
    sc, mag = mr2(xd[:, :2048].contiguous(), yd[:, :2048].contiguous())
    assert rel_l2(torch.stack([sc, mag]).cpu(), g["mr_small"]) < REL_TOL
    mg = losses.stft(xd[:1, :4096].contiguous(), 1024, 120, 600, torch.hann_window(600).to(dev))
    assert rel_l2(mg[:, :8].cpu(), g["stft_mag"]) < REL_TOL
    for tag, kw in (("v1", dict(fs=22050, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80, fmin=0, fmax=11025, log_base=None)),
                    ("default", dict())):
        ms = losses.MelSpectrogram(**kw).to(dev)
        assert torch.equal(ms.melmat.cpu(), g[f"melmat_{tag}"])
        assert rel_l2(ms(xd[:2].contiguous()).cpu(), g[f"mel_{tag}"]) < REL_TOL
        ml = losses.MelSpectrogramLoss(**kw).to(dev)(xd.unsqueeze(1), yd.unsqueeze(1))
        assert rel_l2(ml.reshape(1).cpu(), g[f"mel_loss_{tag}"]) < REL_TOL
    gen = torch.Generator().manual_seed(77)
    mk = lambda: [[torch.randn(2, 4, 50, generator=gen), torch.randn(2, 8, 25, generator=gen), torch.randn(2, 1, 25, generator=gen)] for _ in range(3)]  # noqa: E731
    oh, o = mk(), mk()
    ohd = [[t.to(dev) for t in l] for l in oh]
    od = [[t.to(dev) for t in l] for l in o]
    for lt in ("mse", "hinge"):
        assert rel_l2(losses.GeneratorAdversarialLoss(loss_type=lt)(ohd).reshape(1).cpu(), g[f"gen_adv_{lt}"]) < 1e-5
        r, f = losses.DiscriminatorAdversarialLoss(loss_type=lt)(ohd, od)
        assert rel_l2(torch.stack([r, f]).cpu(), g[f"dis_adv_{lt}"]) < 1e-5
    assert rel_l2(losses.FeatureMatchLoss()(ohd, od).reshape(1).cpu(), g["feat_match"]) < 1e-5
    assert rel_l2(losses.FeatureMatchLoss(False, False, True)(ohd, od).reshape(1).cpu(), g["feat_match_noavg"]) < 1e-5


def test_mr_stft_loss_properties_full_size(dev):
    """BASELINE C3 size (64 x 25600): size-independent properties -- identical signals give
    (sc, mag) ~ (0, 0); the loss is invariant to batch order; scaling both signals by a power of two
    leaves sc unchanged (magnitudes scale exactly) wherever the clamp is inactive."""
    from parallelwavegan_b200 import losses

    mr = losses.MultiResolutionSTFTLoss().to(dev)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(64, 25600, generator=g) - 0.5).to(dev)
    y = (torch.rand(64, 25600, generator=g) - 0.5).to(dev)
    sc, mag = mr(x, x)
    # X and Y come out of the two halves of one complex FFT, so they agree to rounding, not bitwise
    assert float(sc) < 1e-6 and float(mag) < 1e-6
    sc1, mag1 = mr(x, y)
    perm = torch.randperm(64, generator=g).to(dev)
    sc2, mag2 = mr(x[perm].contiguous(), y[perm].contiguous())
    assert abs(float(sc1) - float(sc2)) < 1e-6 * float(sc1) and abs(float(mag1) - float(mag2)) < 1e-6 * float(mag1)
    sc3, _ = mr(4.0 * x, 4.0 * y)
    assert abs(float(sc3) - float(sc1)) < 1e-5 * float(sc1)


def test_style_melgan_discriminator_full_tensors_and_gradients(dev):
    """StyleMelGANDiscriminator (style_melgan.py:243-337): every feature map in full vs the oracle for the same window
    positions, and d(sum of logits)/dx through the random-window crops + PQMF analysis + MelGAN discriminators vs torch
    autograd on the oracle (bare sum: no activation-mask conditioning enters the comparison at the output)."""
    import numpy as np

    from parallelwavegan_b200 import models
    from oracle.ref_ops import fold_weight_norm

    m = models.StyleMelGANDiscriminator()
    spec = [(k, tuple(v.shape)) for k, v in m.state_dict().items() if not k.endswith("_filter")]
    sd = synth.synth_state_dict(spec, 66, 1.4)
    m.load_state_dict(sd, strict=False)
    m = m.to(dev)
    x = synth.randn((2, 1, 5000), 67, 0.5)
    T = x.shape[-1]
    np.random.seed(5)
    starts = [np.random.randint(T - ws) for _ in range(m.repeats) for ws in m.window_sizes]
    np.random.seed(5)
    xd = x.to(dev).requires_grad_(True)
    outs = m(xd)
    w = fold_weight_norm(sd)
    xr = x.clone().requires_grad_(True)
    ref = ref_ops.style_melgan_discriminator(w, xr, starts)
    assert len(outs) == len(ref) == 8
    for i, (o, r) in enumerate(zip(outs, ref)):
        assert len(o) == len(r) == 7
        for j, (a, b) in enumerate(zip(o, r)):
            assert tuple(a.shape) == tuple(b.shape)
            assert rel_l2(a.detach().cpu(), b.detach()) < REL_TOL, (i, j)
    sum(o[-1].sum() for o in outs).backward()
    sum(r[-1].sum() for r in ref).backward()
    assert rel_l2(xd.grad.cpu(), xr.grad) < 5e-3  # LeakyReLU masks inside the towers: conditioning bound, see tests/test_gpu_backward.py
