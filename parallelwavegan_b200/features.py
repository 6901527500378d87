"""Feature extraction on the GPU: ``logmelfilterbank`` of bin/preprocess.py:26-89 on the fused STFT /
mel kernels of the loss path (pwgb_stft_amplitude_forward + pwgb_mel_project_forward), so that analysis-synthesis
(wav -> log-mel -> vocoder) never leaves the device.  Same signature, same (#frames, num_mels) output; the reference's
own test pins this function to ``MelSpectrogram`` at 6 decimals (test/test_mel_loss.py:16-46)."""
import torch

from .capi import PwgbError
from .losses import MelSpectrogram

_CACHE = {}


def logmelfilterbank(audio, sampling_rate, fft_size=1024, hop_size=256, win_length=None, window="hann", num_mels=80, fmin=None,
                     fmax=None, eps=1e-10, log_base=10.0):
    """audio: (T,) or (B, T) float32 CUDA tensor -> (#frames, num_mels) or (B, #frames, num_mels) log-mel features."""
    if not isinstance(audio, torch.Tensor) or not audio.is_cuda:
        raise PwgbError("logmelfilterbank: expected a CUDA tensor (no CPU fallback; the reference's librosa path is the CPU implementation)")
    key = (sampling_rate, fft_size, hop_size, win_length, window, num_mels, fmin, fmax, eps, log_base, audio.device)
    ms = _CACHE.get(key)
    if ms is None:
        ms = MelSpectrogram(fs=sampling_rate, fft_size=fft_size, hop_size=hop_size, win_length=win_length, window=window,
                            num_mels=num_mels, fmin=fmin, fmax=fmax, eps=eps, log_base=log_base).to(audio.device)
        _CACHE[key] = ms
    x = audio.reshape(1, -1) if audio.dim() == 1 else audio
    with torch.no_grad():
        mel = ms(x.float().contiguous())  # (B, num_mels, frames)
    out = mel.transpose(1, 2).contiguous()
    return out[0] if audio.dim() == 1 else out
