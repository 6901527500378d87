"""Boundary helpers mirroring ``parallel_wavegan.utils`` for the hot path: ``load_model`` (utils/utils.py:294-360,
the entry ESPnet / notebooks / ``parallel-wavegan-decode`` use) and the checkpoint dict layout of
``Trainer.save_checkpoint`` / ``load_checkpoint`` (bin/train.py:112-186).  Checkpoints written by the reference load
into the mirror modules unchanged and vice versa (same ``state_dict`` keys, same dict nesting)."""
import os
from packaging.version import Version

import torch
import yaml


def load_model(checkpoint, config=None, stats=None):
    """Load a trained generator (utils/utils.py:294-360).

    checkpoint: path of a ``checkpoint-*.pkl`` (dict with ``model.generator``); config: dict, or None to read
    ``config.yml`` next to the checkpoint; stats: statistics file (``.npy``; ``.h5`` needs h5py) or None to pick up
    ``stats.{npy,h5}`` next to the checkpoint."""
    if config is None:
        with open(os.path.join(os.path.dirname(checkpoint), "config.yml")) as f:
            config = yaml.load(f, Loader=yaml.Loader)
    from . import models

    generator_type = config.get("generator_type", "ParallelWaveGANGenerator")
    if not hasattr(models, generator_type):
        from .capi import PwgbError

        raise PwgbError(f"generator_type={generator_type!r} is not on the B200 hot path (SURVEY.md 8: out of scope)")
    model_class = getattr(models, generator_type)
    # workaround for the reference's typo #295 (utils.py:322-326)
    generator_params = {k.replace("upsample_kernal_sizes", "upsample_kernel_sizes"): v for k, v in config["generator_params"].items()}
    model = model_class(**generator_params)
    model.load_state_dict(torch.load(checkpoint, map_location="cpu")["model"]["generator"])
    if stats is None:
        dirname = os.path.dirname(checkpoint)
        ext = "h5" if config.get("format", "hdf5") == "hdf5" else "npy"
        if os.path.exists(os.path.join(dirname, f"stats.{ext}")):
            stats = os.path.join(dirname, f"stats.{ext}")
    if stats is not None and generator_type != "VQVAE":
        model.register_stats(stats)
    if config["generator_params"]["out_channels"] > 1:
        from .layers import PQMF

        pqmf_params = {}
        if Version(str(config.get("version", "0.1.0"))) <= Version("0.4.2"):
            pqmf_params.update(taps=62, cutoff_ratio=0.15, beta=9.0)  # defaults of versions <= 0.4.2 (utils.py:348-351)
        model.pqmf = PQMF(subbands=config["generator_params"]["out_channels"], **config.get("pqmf_params", pqmf_params))
    return model


def _unwrap(m):
    return m.module if hasattr(m, "module") else m


def save_checkpoint(checkpoint_path, model, optimizer, scheduler=None, steps=0, epochs=0):
    """``Trainer.save_checkpoint`` (train.py:112-146): model / optimizer / scheduler are dicts with the keys
    ``generator`` and ``discriminator`` (DDP wrappers are unwrapped like the reference's ``.module``)."""
    state = {
        "optimizer": {k: optimizer[k].state_dict() for k in ("generator", "discriminator")},
        "scheduler": {k: scheduler[k].state_dict() for k in ("generator", "discriminator")} if scheduler else {"generator": {}, "discriminator": {}},
        "steps": steps,
        "epochs": epochs,
        "model": {k: _unwrap(model[k]).state_dict() for k in ("generator", "discriminator")},
    }
    d = os.path.dirname(checkpoint_path)
    if d and not os.path.exists(d):
        os.makedirs(d)
    torch.save(state, checkpoint_path)


def load_checkpoint(checkpoint_path, model, optimizer=None, scheduler=None, load_only_params=False):
    """``Trainer.load_checkpoint`` (train.py:148-186).  Returns (steps, epochs) (0, 0 with ``load_only_params``)."""
    state = torch.load(checkpoint_path, map_location="cpu")
    _unwrap(model["generator"]).load_state_dict(state["model"]["generator"])
    _unwrap(model["discriminator"]).load_state_dict(state["model"]["discriminator"], strict=False)
    if load_only_params:
        return 0, 0
    for k in ("generator", "discriminator"):
        if optimizer is not None:
            optimizer[k].load_state_dict(state["optimizer"][k])
        if scheduler is not None and state["scheduler"].get(k):
            scheduler[k].load_state_dict(state["scheduler"][k])
    return state["steps"], state["epochs"]
