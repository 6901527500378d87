"""GPU side of the data path: device collater vs the pinned restatement of the reference
Collater (same np.random seed -> the same batch, bit for bit), log-mel extraction vs the oracle, standalone magnitude
losses (losses/stft_loss.py:43-82)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, rel_l2
from oracle import ref_data, ref_ops, synth
from oracle.make_golden_optim import data_items

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    import __graft_entry__

    __graft_entry__.build()
    return torch.device("cuda:0")


def test_device_collater_matches_reference_batches(dev):
    from parallelwavegan_b200 import datasets

    g = np.load(os.path.join(GOLD, "data.npz"))
    items = data_items()
    corpus = datasets.DeviceCorpus(items, dev, hop_size=64)
    idx = list(range(len(items)))
    col = datasets.Collater(batch_max_steps=1100, hop_size=64, aux_context_window=2, use_noise_input=False)
    np.random.seed(11)
    (c,), y = col(corpus, idx)
    assert torch.equal(c.cpu(), torch.from_numpy(g["mel2wav_c"])) and torch.equal(y.cpu(), torch.from_numpy(g["mel2wav_y"]))
    col = datasets.Collater(batch_max_steps=512, hop_size=64, aux_context_window=0, use_noise_input=True)
    np.random.seed(12)
    (z, c), y = col(corpus, idx)
    assert torch.equal(c.cpu(), torch.from_numpy(g["noise_c"])) and torch.equal(y.cpu(), torch.from_numpy(g["noise_y"]))
    assert tuple(z.shape) == tuple(g["noise_z_shape"]) and z.is_cuda and abs(float(z.mean())) < 0.2 and 0.8 < float(z.std()) < 1.2
    audio_only = datasets.DeviceCorpus([x for x, _ in items], dev)
    col = datasets.Collater(batch_max_steps=1500, hop_size=None, aux_context_window=0, use_aux_input=False)
    np.random.seed(13)
    (l, gcond), y = col(audio_only, idx)
    assert l is None and gcond is None and torch.equal(y.cpu(), torch.from_numpy(g["audio_y"]))


def test_device_collater_c5_batch(dev):
    """LJSpeech-like corpus slice at the C5 batch (16 x 8192 samples, 80 mels, hop 256, ctx 0): vs the restatement."""
    from parallelwavegan_b200 import datasets

    rng = np.random.RandomState(5)
    items = [(rng.randn(f * 256).astype(np.float32), rng.randn(f, 80).astype(np.float32)) for f in rng.randint(40, 700, size=24)]
    corpus = datasets.DeviceCorpus(items, dev, hop_size=256)
    idx = list(rng.permutation(24)[:16])
    np.random.seed(99)
    cr, yr = ref_data.collate_mel2wav([items[i] for i in idx], batch_max_steps=8192, hop_size=256, aux_context_window=0)
    np.random.seed(99)
    (c,), y = datasets.Collater(batch_max_steps=8192, hop_size=256, aux_context_window=0)(corpus, idx)
    assert tuple(y.shape) == (16, 1, 8192) and tuple(c.shape) == (16, 80, 32)
    assert torch.equal(c.cpu(), torch.from_numpy(cr)) and torch.equal(y.cpu(), torch.from_numpy(yr))


def test_logmelfilterbank_vs_oracle(dev):
    from parallelwavegan_b200 import features

    x = synth.randn((2, 22050), 41, 0.3)
    for kw in (dict(sampling_rate=22050, fft_size=1024, hop_size=256, win_length=None, num_mels=80, fmin=80, fmax=7600),
               dict(sampling_rate=24000, fft_size=2048, hop_size=300, win_length=1200, num_mels=80, fmin=0, fmax=None)):
        got = features.logmelfilterbank(x.to(dev), window="hann", **kw).cpu().numpy()
        for b in range(2):
            ref = ref_data.logmelfilterbank(x[b].numpy(), **kw)
            assert got[b].shape == ref.shape
            np.testing.assert_almost_equal(got[b], ref, decimal=4)
        one = features.logmelfilterbank(x[0].to(dev), window="hann", **kw)
        assert one.dim() == 2 and torch.equal(one.cpu(), torch.from_numpy(got[0]))


def test_standalone_magnitude_losses(dev):
    from parallelwavegan_b200 import losses

    x = synth.randn((3, 8192), 501, 0.3)
    y = 0.7 * synth.randn((3, 8192), 502, 0.3) + 0.3 * x
    win = torch.hann_window(600)
    xm = losses.stft(x.to(dev), 1024, 120, 600, win.to(dev))
    ym = losses.stft(y.to(dev), 1024, 120, 600, win.to(dev))
    sc = losses.SpectralConvergenceLoss()(xm, ym)
    mag = losses.LogSTFTMagnitudeLoss()(xm, ym)
    xr, yr = ref_ops.stft_mag(x, 1024, 120, 600), ref_ops.stft_mag(y, 1024, 120, 600)
    sc_ref = torch.norm(yr - xr, p="fro") / torch.norm(yr, p="fro")
    mag_ref = torch.nn.functional.l1_loss(torch.log(yr), torch.log(xr))
    assert sc.dim() == 0 and sc.is_cuda
    assert abs(float(sc) - float(sc_ref)) < 1e-4 * float(sc_ref) and abs(float(mag) - float(mag_ref)) < 1e-4 * float(mag_ref)
