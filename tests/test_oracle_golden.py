"""Pin the CPU oracle (oracle/ref_ops.py) against golden vectors produced by the
real reference (oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import ORACLE_TOL, golden_effective_weights, load_golden, rel_l2
from oracle import ref_ops, synth


@pytest.mark.parametrize("name", ["hifigan_small", "hifigan_v1", "hifigan_odd", "hifigan_causal"])
def test_hifigan_generator(name):
    meta, g = load_golden(name)
    w = golden_effective_weights(meta)
    kw = meta["kwargs"]
    cfg = dict(kw, negative_slope=kw["nonlinear_activation_params"]["negative_slope"])
    c = synth.randn(meta["c_shape"], meta["c_seed"])
    y = ref_ops.hifigan_generator(w, c, cfg)
    assert y.shape == g["y"].shape
    assert rel_l2(y, g["y"]) < ORACLE_TOL
    assert rel_l2(y[0].t(), g["y_inf"]) < ORACLE_TOL


@pytest.mark.parametrize("name", ["mb_melgan_v2", "melgan_small", "melgan_causal"])
def test_melgan_generator(name):
    meta, g = load_golden(name)
    w = golden_effective_weights(meta)
    cfg = dict(meta["kwargs"], negative_slope=0.2)
    c = synth.randn(meta["c_shape"], meta["c_seed"])
    y = ref_ops.melgan_generator(w, c, cfg)
    assert rel_l2(y, g["y"]) < ORACLE_TOL
    if meta.get("pqmf_subbands"):
        _, syn = ref_ops.pqmf_filters(meta["pqmf_subbands"])
        yp = ref_ops.pqmf_synthesis(y, syn)
        assert rel_l2(yp, g["y_pqmf"]) < ORACLE_TOL
        assert rel_l2(yp[0].t(), g["y_inf"]) < ORACLE_TOL


@pytest.mark.parametrize("name", ["style_melgan_v1", "style_melgan_small"])
def test_style_melgan_generator(name):
    meta, g = load_golden(name)
    w = golden_effective_weights(meta)
    kw = meta["kwargs"]
    cfg = dict(kw, noise_upsample_negative_slope=kw["noise_upsample_activation_params"]["negative_slope"])
    c = synth.randn(meta["c_shape"], meta["c_seed"])
    z = synth.randn(meta["z_shape"], meta["z_seed"])
    y = ref_ops.style_melgan_generator(w, c, z, cfg)
    assert y.shape == g["y"].shape
    # 9 instance-normalised blocks amplify fp32 reassociation differences (module vs functional ATen calls):
    # 2.4e-5 end to end, 5e-7 per block
    assert rel_l2(y, g["y"]) < 5 * ORACLE_TOL
    # first TADE residual block in isolation (layers/tade_res_block.py:135-160)
    x1, c1 = ref_ops.tade_res_block(w, "blocks.0", torch.from_numpy(g["x0"]) if not isinstance(g["x0"], torch.Tensor) else g["x0"], c,
                                    kw["kernel_size"], kw["dilation"], kw["upsample_scales"][0], kw["gated_function"])
    assert rel_l2(x1, g["x1"]) < ORACLE_TOL and rel_l2(c1, g["c1"]) < ORACLE_TOL


@pytest.mark.parametrize("name", ["pwg_v1", "pwg_small"])
def test_pwg_generator(name):
    meta, g = load_golden(name)
    w = golden_effective_weights(meta)
    kw = meta["kwargs"]
    cfg = dict(kw, upsample_scales=kw["upsample_params"]["upsample_scales"])
    c = synth.randn(meta["c_shape"], meta["c_seed"])
    z = synth.randn(meta["z_shape"], meta["z_seed"])
    c_up = ref_ops.pwg_upsample_net(w, c, cfg)
    assert rel_l2(c_up[:, :, :512], g["c_up"]) < ORACLE_TOL
    x0 = F.conv1d(z, w["first_conv.weight"], w["first_conv.bias"])
    x1, s1 = ref_ops.wavenet_residual_block(w, "conv_layers.0", x0, c_up, 1, kw["kernel_size"])
    assert rel_l2(x1[:, :, :256], g["x1"]) < ORACLE_TOL
    assert rel_l2(s1[:, :, :256], g["s1"]) < ORACLE_TOL
    y = ref_ops.pwg_generator(w, z, c, cfg)
    assert rel_l2(y, g["y"]) < ORACLE_TOL


@pytest.mark.parametrize("n", [2, 3, 4, 8])
def test_pqmf(n):
    meta, g = load_golden(f"pqmf_{n}")
    an, sy = ref_ops.pqmf_filters(n)
    # filter design is float64 numpy -> float32: must be bit-identical to the reference buffers
    assert torch.equal(an, g["analysis_filter"])
    assert torch.equal(sy, g["synthesis_filter"])
    x = synth.randn(meta["x_shape"], meta["x_seed"])
    a = ref_ops.pqmf_analysis(x, an)
    assert rel_l2(a, g["analysis"]) < ORACLE_TOL
    s = ref_ops.pqmf_synthesis(g["analysis"], sy)
    assert rel_l2(s, g["synthesis"]) < ORACLE_TOL


def test_conv_cases_numpy_definition():
    """The float64 direct-loop definitions agree with ATen (golden from the reference's torch)."""
    meta, g = load_golden("conv_cases")
    for case in meta["cases"]:
        i = case["idx"]
        x = synth.randn((2, case["cin"], case["T"]), 200 + i)
        b = synth.randn((case["cout"],), 400 + i, 0.1)
        if case["op"] == "conv1d":
            w = synth.randn((case["cout"], case["cin"] // case["groups"], case["k"]), 300 + i, 0.3)
            y = ref_ops.np_conv1d(x, w, b, case["stride"], case["padding"], case["dilation"], case["groups"])
        else:
            s = case["stride"]
            w = synth.randn((case["cin"], case["cout"], case["k"]), 300 + i, 0.3)
            y = ref_ops.np_conv_transpose1d(x, w, b, s, s // 2 + s % 2, s % 2)
        assert y.shape == tuple(g[f"y{i}"].shape)
        assert rel_l2(y, g[f"y{i}"]) < ORACLE_TOL


def test_kaiser_matches_scipy():
    import scipy.signal.windows

    np.testing.assert_allclose(ref_ops.kaiser_window(63, 9.0), scipy.signal.windows.kaiser(63, 9.0), rtol=0, atol=1e-15)


def test_mel_filterbank_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for fmin, fmax in ((0.0, 11025.0), (80.0, 7600.0)):
        ours = ref_ops.slaney_mel_filterbank(22050, 1024, 80, fmin, fmax)
        theirs = ta.functional.melscale_fbanks(513, fmin, fmax, 80, 22050, norm="slaney", mel_scale="slaney").T.numpy()
        assert np.abs(ours - theirs).max() < 5e-7
