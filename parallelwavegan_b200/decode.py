"""Decode driver -- the GPU side of ``parallel-wavegan-decode`` (bin/decode.py:214-243).

The reference loops utterances one at a time through ``model.inference`` (host normalisation, one forward of
~80 launches, synchronous float D2H, ``sf.write(..., "PCM_16")``).  Here:

* utterances are sharded over ranks (``i mod world``; no collective) and batched by length;
* feature normalisation ``(c - mean) / scale`` (hifigan.py:264-265), the (T, C) -> (C, T) transpose and the edge
  padding (``ReplicationPad1d(aux_context_window)`` for Parallel WaveGAN) are one kernel per utterance writing straight
  into the batch slot (``pwgb_prep_features``);
* the generator forward of every distinct batch shape is captured once into a CUDA graph and replayed;
* the waveform is quantised to PCM16 on the GPU (``pwgb_pcm16_forward``, libsndfile's rule) and leaves the device
  as int16 through an asynchronous copy into pinned host memory -- an utterance crosses PCIe once each way.

``exact=True`` (default) batches only utterances of EQUAL length, so every waveform equals ``model.inference`` of that
utterance alone; ``exact=False`` pads a length bucket with zero frames (after normalisation) to a multiple of
``bucket_frames`` and crops -- samples within one receptive field of an utterance's end then see the padded tail
instead of the conv's zero padding."""
import ctypes as C

import numpy as np
import torch

from . import capi, ops, sharding
from .capi import PAD_REPLICATE, PAD_ZERO, PwgbError


def _stream():
    return ops._stream()


class GraphedGenerator:
    """Replays ``fn(*static_inputs)`` from a CUDA graph for every distinct tuple of input shapes.

    The first call with new shapes runs two eager warm-ups on a side stream (packs weights, sets kernel
    attributes, fills every per-shape cache), then captures; later calls copy the inputs into the static
    tensors and replay.  The graph holds the packed weight images: call ``reset()`` after changing weights."""

    def __init__(self, model, max_graphs=16, fn=None):
        self.model = model.eval() if hasattr(model, "eval") else model
        self.fn = fn if fn is not None else self.model
        self.max_graphs = max_graphs
        self._graphs = {}

    def reset(self):
        self._graphs.clear()

    @torch.no_grad()
    def __call__(self, *inputs):
        key = tuple(tuple(t.shape) for t in inputs)
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.max_graphs:
                return self.fn(*inputs)
            static_in = [t.clone() for t in inputs]
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.fn(*static_in)
            ent = (graph, static_in, static_out)
            self._graphs[key] = ent
        graph, static_in, static_out = ent
        for s, t in zip(static_in, inputs):
            s.copy_(t)
        graph.replay()
        return static_out


def pcm16(y):
    """float waveform (any shape, CUDA) -> int16 tensor of the same shape (libsndfile's PCM_16 rule)."""
    if not y.is_cuda or y.dtype != torch.float32:
        raise PwgbError("pcm16: expected a float32 CUDA tensor (no CPU fallback)")
    y = y.contiguous()
    out = torch.empty(y.shape, device=y.device, dtype=torch.int16)
    rc = capi.lib().pwgb_pcm16_forward(C.c_void_p(y.data_ptr()), C.c_void_p(out.data_ptr()), y.numel(), _stream())
    capi.check(rc, "pwgb_pcm16_forward")
    return out


def prep_features(c, out, mean=None, scale=None, pad_left=0, replicate=False):
    """c: (T, C) features on the device -> out: (C, T_out) slot (normalised, transposed, edge-padded)."""
    T, Cc = c.shape
    rc = capi.lib().pwgb_prep_features(C.c_void_p(c.data_ptr()), C.c_void_p(mean.data_ptr()) if mean is not None else None,
                                       C.c_void_p(scale.data_ptr()) if scale is not None else None, C.c_void_p(out.data_ptr()), T, Cc,
                                       int(pad_left), out.shape[-1], PAD_REPLICATE if replicate else PAD_ZERO, _stream())
    capi.check(rc, "pwgb_prep_features")
    return out


class Decoder:
    def __init__(self, model, use_graphs=True, max_batch=16, exact=True, bucket_frames=32, rank=0, world=1, seed=None):
        from . import models

        self.model = model.eval()
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise PwgbError("Decoder: the model must live on a CUDA device (no CPU fallback)")
        self.kind = ("pwg" if isinstance(model, models.ParallelWaveGANGenerator) else
                     "style" if isinstance(model, models.StyleMelGANGenerator) else "mel2wav")
        self.max_batch, self.exact, self.bucket = int(max_batch), bool(exact), int(bucket_frames)
        self.rank, self.world = rank, world
        self.gen = torch.Generator(device=self.dev)
        if seed is not None:
            self.gen.manual_seed(int(seed))
        pq = getattr(model, "pqmf", None)

        def fwd(*a):
            y = model(*a)
            return pq.synthesis(y) if pq is not None else y

        self.runner = GraphedGenerator(model, fn=fwd) if use_graphs else fwd
        self.hop = int(getattr(model, "upsample_factor", 0)) or None
        self.ctx = int(getattr(model, "aux_context_window", 0)) if self.kind == "pwg" else 0

    def _groups(self, lengths, idxs):
        groups = {}
        for i in idxs:
            L = lengths[i]
            key = L if self.exact else -(-L // self.bucket) * self.bucket
            groups.setdefault(key, []).append(i)
        for key in sorted(groups):
            g = groups[key]
            for k in range(0, len(g), self.max_batch):
                yield key, g[k : k + self.max_batch]

    @torch.no_grad()
    def decode(self, mels, normalize_before=False, to_pcm16=True, noises=None):
        """mels: list of (T_i, C) arrays / tensors.  Returns {index: waveform}: int16 numpy arrays of shape (samples,)
        (``to_pcm16``) or float32 device tensors (samples, out_channels).  ``noises`` (Parallel WaveGAN only): optional
        {index: (samples, 1)} noise for reproducibility (``inference(c, x)``)."""
        m = self.model
        idxs = list(sharding.partition(len(mels), self.rank, self.world))
        if self.kind == "style":  # batch-1 inference with its own noise / padding protocol (style_melgan.py:226-262)
            out = {}
            for i in idxs:
                y = m.inference(torch.as_tensor(mels[i], dtype=torch.float32, device=self.dev), normalize_before=normalize_before)
                out[i] = pcm16(y[:, 0]).cpu().numpy() if to_pcm16 else y
            return out
        mean = m.mean if normalize_before else None
        scale = m.scale if normalize_before else None
        host = [torch.as_tensor(np.asarray(x, dtype=np.float32) if not isinstance(x, torch.Tensor) else x, dtype=torch.float32) for x in (mels[i] for i in idxs)]
        by_idx = dict(zip(idxs, host))
        lengths = {i: by_idx[i].shape[0] for i in idxs}
        out, pending = {}, []
        for Tb, grp in self._groups(lengths, idxs):
            B, Cc = len(grp), by_idx[grp[0]].shape[1]
            cin = torch.empty((B, Cc, Tb + 2 * self.ctx), device=self.dev, dtype=torch.float32)
            for b, i in enumerate(grp):
                src = by_idx[i]
                src = src.pin_memory() if not src.is_cuda and not src.is_pinned() else src
                prep_features(src.to(self.dev, non_blocking=True).contiguous(), cin[b], mean, scale, pad_left=self.ctx, replicate=self.ctx > 0)
            if self.kind == "pwg":
                hop = m.upsample_factor
                z = torch.randn((B, 1, Tb * hop), device=self.dev, generator=self.gen)
                if noises:
                    for b, i in enumerate(grp):
                        if i in noises:
                            nz = torch.as_tensor(noises[i], dtype=torch.float32).to(self.dev).reshape(-1)
                            z[b, 0, : nz.numel()] = nz
                y = self.runner(z, cin)
            else:
                y = self.runner(cin)
            hop = y.shape[-1] // Tb
            q = pcm16(y) if to_pcm16 else None
            for b, i in enumerate(grp):
                n = lengths[i] * hop
                if to_pcm16:
                    hbuf = torch.empty(n, dtype=torch.int16).pin_memory()
                    hbuf.copy_(q[b, 0, :n], non_blocking=True)
                    pending.append((i, hbuf))
                else:
                    out[i] = y[b, :, :n].transpose(1, 0).clone()
        if pending:
            torch.cuda.current_stream().synchronize()
            for i, hbuf in pending:
                out[i] = hbuf.numpy()
        return out


@torch.no_grad()
def decode_utterances(model, mels, rank=0, world=1, normalize_before=False, use_graphs=True):
    """Decode this rank's share (i mod world == rank) of ``mels`` (list of (T', C) tensors / arrays), every utterance
    bit-identical to ``model.inference`` of it alone.  Returns {index: waveform (T, out_channels) on the device}."""
    return Decoder(model, use_graphs=use_graphs, rank=rank, world=world).decode(mels, normalize_before=normalize_before, to_pcm16=False)
