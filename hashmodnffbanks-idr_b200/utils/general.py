"""Glue helpers with the reference's names (utils/general.py:9-50)."""
import importlib

import torch


def get_class(kls):
    """'pkg.mod.Class' -> class object (the reference resolves train.model_class this way)."""
    module, _, name = kls.rpartition('.')
    return getattr(importlib.import_module(module), name)


def split_input(model_input, total_pixels, n_pixels=10000):
    """Splits uv / object_mask into chunks of n_pixels rays (eval-time rendering)."""
    device = model_input['uv'].device
    chunks = []
    for idx in torch.split(torch.arange(total_pixels, device=device), n_pixels, dim=0):
        part = dict(model_input)
        part['uv'] = torch.index_select(model_input['uv'], 1, idx)
        part['object_mask'] = torch.index_select(model_input['object_mask'], 1, idx)
        chunks.append(part)
    return chunks


def merge_output(res, total_pixels, batch_size):
    merged = {}
    for key, first in res[0].items():
        if first is None:
            continue
        if first.dim() == 1:
            merged[key] = torch.cat([r[key].reshape(batch_size, -1, 1) for r in res], 1).reshape(batch_size * total_pixels)
        else:
            merged[key] = torch.cat([r[key].reshape(batch_size, -1, r[key].shape[-1]) for r in res], 1) \
                .reshape(batch_size * total_pixels, -1)
    return merged


def render_image(model, model_input, total_pixels, n_pixels=10000, keys=("rgb_values",)):
    """Full-image rendering as the reference's evaluation loop does it (evaluation/eval.py:150-160): the image's
    pixels go through `model` (eval mode) in splits of `n_pixels` rays and the per-split outputs are merged back
    to [batch * total_pixels, C].  Every split runs the device-driven tracer and the eval branch of
    IDRNetwork.forward; nothing but the merged result leaves the device."""
    was_training = model.training
    model.eval()
    try:
        res = []
        for part in split_input(model_input, total_pixels, n_pixels):
            out = model(part)
            res.append({k: out[k].detach() for k in keys})
    finally:
        model.train(was_training)
    return merge_output(res, total_pixels, model_input['uv'].shape[0])
