"""Input pipeline on the GPU: a device-resident corpus + the training collater.

``DeviceCorpus`` keeps every utterance of a training set in HBM (audio concatenated in one tensor, feature
matrices concatenated along frames); ``Collater`` mirrors ``parallel_wavegan.bin.train.Collater`` (train.py:646-925):
same constructor arguments, same filtering of short utterances, the SAME host RNG call sequence for the crop
positions (one ``np.random.randint(start_offset, length + end_offset)`` per kept item, train.py:739-744), so a given
``np.random.seed`` yields the reference's batch bit for bit; the gather itself is one kernel (``pwgb_collate_crop``)
and the noise ``z`` of the Parallel WaveGAN recipes is drawn on the device."""
import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import PwgbError


class DeviceCorpus:
    """items: list of (audio (T,), feats (frames, C)) numpy arrays / tensors [mel2wav case] or of audio arrays
    [audio-only case].  ``hop_size`` applies ``Collater._adjust_length`` (train.py:868-893): audio shorter than
    ``frames * hop_size`` is edge-padded; any other mismatch is an error, like the reference's assert."""

    def __init__(self, items, device, hop_size=None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise PwgbError("DeviceCorpus lives on a CUDA device (no CPU fallback)")
        auds, fts, self.x_start, self.c_start, self.x_len, self.c_len = [], [], [], [], [], []
        xo = co = 0
        self.has_feats = len(items) > 0 and isinstance(items[0], (tuple, list))
        for it in items:
            x, c = (it[0], it[1]) if self.has_feats else (it, None)
            x = np.asarray(x, dtype=np.float32).reshape(-1)
            if c is not None:
                c = np.asarray(c, dtype=np.float32)
                if hop_size is not None:
                    if len(x) < len(c) * hop_size:
                        x = np.pad(x, (0, len(c) * hop_size - len(x)), mode="edge")
                    if len(x) != len(c) * hop_size:
                        raise PwgbError(f"audio / feature length mismatch: {len(x)} != {len(c)} * {hop_size}")
                fts.append(torch.from_numpy(c))
                self.c_start.append(co)
                self.c_len.append(len(c))
                co += len(c)
            auds.append(torch.from_numpy(x))
            self.x_start.append(xo)
            self.x_len.append(len(x))
            xo += len(x)
        self.audio = torch.cat(auds).to(self.device) if auds else torch.empty(0, device=self.device)
        self.feats = torch.cat(fts).contiguous().to(self.device) if fts else None
        self.channels = self.feats.shape[1] if self.feats is not None else 0

    def __len__(self):
        return len(self.x_start)


class Collater:
    """``Collater`` of bin/train.py:646-925 for the mel2wav and audio-only cases, on the device."""

    def __init__(self, batch_max_steps=20480, hop_size=256, aux_context_window=2, use_noise_input=False, use_aux_input=True):
        if hop_size is not None:
            if batch_max_steps % hop_size != 0:
                batch_max_steps += -(batch_max_steps % hop_size)
            assert batch_max_steps % hop_size == 0
            self.hop_size = hop_size
            self.batch_max_frames = batch_max_steps // hop_size
        self.batch_max_steps = batch_max_steps
        self.aux_context_window = aux_context_window
        self.use_noise_input = use_noise_input
        self.use_aux_input = use_aux_input
        if not use_aux_input:
            assert not use_noise_input, "Not supported."
        if use_aux_input:
            self.start_offset = aux_context_window
            self.end_offset = -(self.batch_max_frames + aux_context_window)
            self.mel_threshold = self.batch_max_frames + 2 * aux_context_window
        else:
            self.start_offset = 0
            self.end_offset = -self.batch_max_steps
            self.audio_threshold = self.batch_max_steps
        self.generator = None  # optional torch.Generator (device) for the noise input

    def __call__(self, corpus, indices):
        """indices: the utterance ids of this batch (what the sampler yields).  Returns ``(inputs, y)`` exactly like
        the reference: inputs = (z, c) / (c,) [mel2wav] with c (B, C, frames + 2*ctx) and y (B, 1, T)."""
        dev = corpus.device
        if self.use_aux_input:
            if not corpus.has_feats:
                raise PwgbError("Collater(use_aux_input=True) needs a corpus with features")
            keep = [i for i in indices if corpus.c_len[i] > self.mel_threshold]  # train.py:727-729
            starts = np.array([np.random.randint(self.start_offset, corpus.c_len[i] + self.end_offset) for i in keep], dtype=np.int64)
            x_off = np.array([corpus.x_start[i] for i in keep], dtype=np.int64) + starts * self.hop_size
            c_off = np.array([corpus.c_start[i] for i in keep], dtype=np.int64) + starts - self.aux_context_window
            F = self.batch_max_frames + 2 * self.aux_context_window
        else:
            keep = [i for i in indices if corpus.x_len[i] >= self.audio_threshold]  # train.py:840-843
            starts = np.array([np.random.randint(self.start_offset, corpus.x_len[i] + self.end_offset) for i in keep], dtype=np.int64)
            x_off = np.array([corpus.x_start[i] for i in keep], dtype=np.int64) + starts
            c_off, F = None, 0
        B, T = len(keep), self.batch_max_steps
        y = torch.empty((B, 1, T), device=dev, dtype=torch.float32)
        xo = torch.from_numpy(x_off).pin_memory().to(dev, non_blocking=True)
        c = co = None
        if c_off is not None:
            co = torch.from_numpy(c_off).pin_memory().to(dev, non_blocking=True)
            c = torch.empty((B, corpus.channels, F), device=dev, dtype=torch.float32)
        vp = C.c_void_p
        rc = capi.lib().pwgb_collate_crop(vp(corpus.audio.data_ptr()), vp(xo.data_ptr()), vp(corpus.feats.data_ptr()) if c is not None else None,
                                          vp(co.data_ptr()) if co is not None else None, vp(y.data_ptr()), vp(c.data_ptr()) if c is not None else None,
                                          B, T, corpus.channels, F, vp(torch.cuda.current_stream().cuda_stream))
        capi.check(rc, "pwgb_collate_crop")
        if not self.use_aux_input:
            return (None, None), y  # the reference returns (l_batch, g_batch) = (None, None) here (train.py:861-876)
        inputs = (c,)
        if self.use_noise_input:
            z = torch.randn(y.shape, device=dev, generator=self.generator)  # train.py:790-792, drawn on the device
            inputs = (z,) + inputs
        return inputs, y
