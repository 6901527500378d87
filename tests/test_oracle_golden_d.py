"""Pin the discriminator / loss parts of the oracle against golden vectors from the real reference."""
import pytest
import torch

from helpers import ORACLE_TOL, check_fingerprint, golden_weights, load_golden, rel_l2
from oracle import ref_ops, synth


def _eff(meta):
    return ref_ops.fold_spectral_norm_eval(ref_ops.fold_weight_norm(golden_weights(meta)))


def test_hifigan_msmpd():
    meta, g = load_golden("hifigan_msmpd_v1")
    x = synth.randn(meta["x_shape"], meta["x_seed"], meta["x_scale"])
    outs = ref_ops.hifigan_msmpd(_eff(meta), x)
    check_fingerprint(outs, meta, g, 5e-5)
    for i, o in enumerate(outs):
        assert rel_l2(o[-1], g[f"final{i}"]) < 5e-5


def test_melgan_msd():
    meta, g = load_golden("melgan_msd")
    x = synth.randn(meta["x_shape"], meta["x_seed"], meta["x_scale"])
    outs = ref_ops.melgan_msd(_eff(meta), x, downsample_scales=(4, 4, 4), max_ch=512)
    check_fingerprint(outs, meta, g, 5e-5)


def test_style_melgan_discriminator():
    """Random-window discriminator (style_melgan.py:243-337): same np.random seed as the fixture -> same windows."""
    import numpy as np

    meta, g = load_golden("style_melgan_disc")
    x = synth.randn(meta["x_shape"], meta["x_seed"], meta["x_scale"])
    T = x.shape[-1]
    np.random.seed(meta["np_seed"])
    starts = [np.random.randint(T - ws) for _ in range(2) for ws in (512, 1024, 2048, 4096)]
    outs = ref_ops.style_melgan_discriminator(_eff(meta), x, starts)
    check_fingerprint(outs, meta, g, 5e-5)
    for i, o in enumerate(outs):
        assert rel_l2(o[-1], g[f"final{i}"]) < 5e-5


def test_pwg_discriminator():
    meta, g = load_golden("pwg_disc")
    x = synth.randn(meta["x_shape"], meta["x_seed"], meta["x_scale"])
    y = ref_ops.pwg_discriminator(_eff(meta), x)
    assert rel_l2(y, g["final0"]) < ORACLE_TOL


def test_losses():
    _, g = load_golden("losses")
    x = synth.randn((3, 8192), 501, 0.3)
    y = synth.randn((3, 8192), 502, 0.3)
    y = 0.7 * y + 0.3 * x
    sc, mag = ref_ops.mr_stft_loss(x, y)
    assert rel_l2(torch.stack([sc, mag]), g["mr_default"]) < 1e-5
    sc, mag = ref_ops.mr_stft_loss(x.view(1, 3, -1), y.view(1, 3, -1))
    assert rel_l2(torch.stack([sc, mag]), g["mr_3d"]) < 1e-5
    sc, mag = ref_ops.mr_stft_loss(x[:, :2048], y[:, :2048], (64, 128, 256), (16, 32, 64), (64, 128, 256))
    assert rel_l2(torch.stack([sc, mag]), g["mr_small"]) < 1e-5
    assert rel_l2(ref_ops.stft_mag(x[:1, :4096], 1024, 120, 600)[:, :8], g["stft_mag"]) < 1e-5
    for tag, kw, fr in (("v1", dict(log_base=None), (0, 11025)), ("default", dict(), (80, 7600))):
        melmat = torch.from_numpy(ref_ops.slaney_mel_filterbank(22050, 1024, 80, *fr).T.copy())
        assert torch.equal(melmat, g[f"melmat_{tag}"])
        assert rel_l2(ref_ops.mel_spectrogram(x[:2], melmat, **kw), g[f"mel_{tag}"]) < 1e-5
        assert rel_l2(ref_ops.mel_loss(x.unsqueeze(1), y.unsqueeze(1), melmat, **kw).reshape(1), g[f"mel_loss_{tag}"]) < 1e-5
    gen = torch.Generator().manual_seed(77)
    mk = lambda: [[torch.randn(2, 4, 50, generator=gen), torch.randn(2, 8, 25, generator=gen), torch.randn(2, 1, 25, generator=gen)] for _ in range(3)]  # noqa: E731
    outs_hat, outs = mk(), mk()
    for lt in ("mse", "hinge"):
        assert rel_l2(ref_ops.generator_adv_loss(outs_hat, lt).reshape(1), g[f"gen_adv_{lt}"]) < 1e-6
        assert rel_l2(torch.stack(ref_ops.discriminator_adv_loss(outs_hat, outs, lt)), g[f"dis_adv_{lt}"]) < 1e-6
    assert rel_l2(ref_ops.feature_match_loss(outs_hat, outs).reshape(1), g["feat_match"]) < 1e-6
    assert rel_l2(ref_ops.feature_match_loss(outs_hat, outs, False, False, True).reshape(1), g["feat_match_noavg"]) < 1e-6
