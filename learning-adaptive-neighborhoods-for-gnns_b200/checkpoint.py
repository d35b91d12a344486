"""Checkpoint save / load in the reference's format: ``train_small_graphs.save_checkpoint``
writes ``{"args", "epoch", "model_state_dict", "optimizer_state_dict"}`` (train_small_graphs.py:210-220) and
``test_best`` reloads ``["model_state_dict"]`` (329-336); the upstream GCNII scripts save a bare ``state_dict``
(full-supervised.py:128).  Parameter names of the drop-in modules equal the reference's, so files move both ways."""
from __future__ import annotations

import torch


def save_checkpoint(fn, args, epoch, model, optimizer=None, lr_scheduler=None):
    torch.save({
        "args": dict(vars(args)) if args is not None and not isinstance(args, dict) else args,
        "epoch": epoch,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict() if optimizer is not None else None,
    }, fn)


def load_checkpoint(fn, model, optimizer=None, map_location=None, strict=True):
    """Loads either format; returns the stored epoch (None for a bare state_dict)."""
    ckpt = torch.load(fn, map_location=map_location, weights_only=False)
    if isinstance(ckpt, dict) and "model_state_dict" in ckpt:
        model.load_state_dict(ckpt["model_state_dict"], strict=strict)
        if optimizer is not None and ckpt.get("optimizer_state_dict") is not None:
            optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        return ckpt.get("epoch")
    model.load_state_dict(ckpt, strict=strict)
    return None
