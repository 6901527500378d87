"""CPU restatement of the data side of the hot path (TEST INFRASTRUCTURE -- only tests/, smoke() and bench.py's
CPU legs may import oracle/).

* ``collate_mel2wav`` / ``collate_audio`` -- ``Collater.__call__`` of bin/train.py:711-876 (random crop of a batch),
  pinned against the real reference by tests/golden/data.npz (oracle/make_golden_optim.py).
* ``logmelfilterbank`` -- bin/preprocess.py:26-89.  It calls ``librosa.stft`` / ``librosa.filters.mel``; librosa
  (setup.py:29, ``librosa>=0.8.0``, unpinned) is absent from the reference tree and from this image, so the
  restatement follows librosa's published algorithm (centered frames, reflect padding, periodic Hann padded to
  n_fft, rFFT; Slaney mel basis = ref_ops.slaney_mel_filterbank) -- PARITY UNPINNED for the librosa part; the
  reference's own test (test/test_mel_loss.py:16-46) pins it to ``MelSpectrogram`` at 6 decimals, which the
  tests reproduce against ref_ops.mel_spectrogram.
"""
import numpy as np

from oracle.ref_ops import slaney_mel_filterbank


def collate_mel2wav(batch, batch_max_steps=20480, hop_size=256, aux_context_window=2):
    """batch: list of (x (T,), c (frames, C)).  Returns (c_batch (B, C, F), y_batch (B, 1, T)) float32;
    consumes np.random exactly like the reference (one randint per kept item, train.py:739-744)."""
    if batch_max_steps % hop_size != 0:
        batch_max_steps += -(batch_max_steps % hop_size)
    frames = batch_max_steps // hop_size
    start_offset, end_offset = aux_context_window, -(frames + aux_context_window)
    thr = frames + 2 * aux_context_window
    kept = []
    for x, c in batch:
        if len(c) > thr:  # train.py:727-729
            if len(x) < len(c) * hop_size:  # _adjust_length, train.py:886-887
                x = np.pad(x, (0, len(c) * hop_size - len(x)), mode="edge")
            assert len(x) == len(c) * hop_size
            kept.append((x, c))
    starts = np.array([np.random.randint(start_offset, len(c) + end_offset) for _, c in kept])
    y = np.array([x[s * hop_size : s * hop_size + batch_max_steps] for (x, _), s in zip(kept, starts)], dtype=np.float32)
    cb = np.array([c[s - aux_context_window : s + frames + aux_context_window] for (_, c), s in zip(kept, starts)], dtype=np.float32)
    return cb.transpose(0, 2, 1), y[:, None, :]


def collate_audio(batch, batch_max_steps=20480):
    """Audio-only case (train.py:838-858)."""
    kept = [x for x in batch if len(x) >= batch_max_steps]
    starts = np.array([np.random.randint(0, len(x) - batch_max_steps) for x in kept])
    return np.array([x[s : s + batch_max_steps] for x, s in zip(kept, starts)], dtype=np.float32)[:, None, :]


def logmelfilterbank(audio, sampling_rate, fft_size=1024, hop_size=256, win_length=None, num_mels=80, fmin=None, fmax=None,
                     eps=1e-10, log_base=10.0):
    """preprocess.py:26-89 with window="hann".  Returns (#frames, num_mels) float32."""
    audio = np.asarray(audio, dtype=np.float32)
    win_length = fft_size if win_length is None else win_length
    n = np.arange(win_length)
    win = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win_length)  # scipy get_window("hann", fftbins=True)
    left = (fft_size - win_length) // 2
    w = np.zeros(fft_size)
    w[left : left + win_length] = win
    xp = np.pad(audio, fft_size // 2, mode="reflect")
    n_frames = 1 + len(audio) // hop_size
    idx = np.arange(fft_size)[None, :] + hop_size * np.arange(n_frames)[:, None]
    spc = np.abs(np.fft.rfft((xp[idx] * w).astype(np.float32), axis=-1)).astype(np.float32)  # librosa keeps complex64
    fmin = 0 if fmin is None else fmin
    fmax = sampling_rate / 2 if fmax is None else fmax
    mel_basis = slaney_mel_filterbank(sampling_rate, fft_size, num_mels, fmin, fmax)
    mel = np.maximum(eps, np.dot(spc, mel_basis.T))
    if log_base is None:
        return np.log(mel)
    if log_base == 10.0:
        return np.log10(mel)
    if log_base == 2.0:
        return np.log2(mel)
    raise ValueError(f"{log_base} is not supported.")
