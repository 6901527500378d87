"""Pins the data-side restatements (oracle/ref_data.py): the Collater restatement against batches produced by the REAL
reference Collater (tests/golden/data.npz, oracle/make_golden_optim.py), logmelfilterbank against the oracle's
MelSpectrogram restatement at the reference's own tolerance (test/test_mel_loss.py:16-46: 6 decimals)."""
import os

import numpy as np
import torch

from helpers import GOLD
from oracle import ref_data, ref_ops, synth
from oracle.make_golden_optim import data_items


def test_collater_restatement_matches_reference():
    g = np.load(os.path.join(GOLD, "data.npz"))
    items = data_items()
    np.random.seed(11)
    c, y = ref_data.collate_mel2wav(items, batch_max_steps=1100, hop_size=64, aux_context_window=2)
    assert np.array_equal(c, g["mel2wav_c"]) and np.array_equal(y, g["mel2wav_y"])
    np.random.seed(12)
    c, y = ref_data.collate_mel2wav(items, batch_max_steps=512, hop_size=64, aux_context_window=0)
    assert np.array_equal(c, g["noise_c"]) and np.array_equal(y, g["noise_y"]) and tuple(g["noise_z_shape"]) == y.shape
    np.random.seed(13)
    y = ref_data.collate_audio([x for x, _ in items], batch_max_steps=1500)
    assert np.array_equal(y, g["audio_y"])


def test_logmelfilterbank_restatement_matches_mel_spectrogram():
    x = synth.randn((8000,), 77, 0.3)
    for kw in (dict(sampling_rate=22050, fft_size=1024, hop_size=256, win_length=None, num_mels=80, fmin=80, fmax=7600),
               dict(sampling_rate=24000, fft_size=2048, hop_size=300, win_length=1200, num_mels=80, fmin=0, fmax=None)):
        a = ref_data.logmelfilterbank(x.numpy(), **kw)
        sr, nfft = kw["sampling_rate"], kw["fft_size"]
        melmat = torch.from_numpy(ref_ops.slaney_mel_filterbank(sr, nfft, 80, kw["fmin"] or 0, kw["fmax"] or sr / 2)).t()
        b = ref_ops.mel_spectrogram(x[None], melmat, nfft, kw["hop_size"], kw["win_length"], 1e-10, 10.0)[0].t().numpy()
        assert a.shape == b.shape
        np.testing.assert_almost_equal(a, b, decimal=4)  # fp32 FFT summation order; the reference test uses 6 on its own pair
