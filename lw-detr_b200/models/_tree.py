"""Builds the nested nn.Module tree that carries the reference's state_dict names from the flat table in
b200/spec.py (so `load_state_dict(strict=True)` of a reference checkpoint works)."""
import math

import torch
from torch import nn


class Holder(nn.Module):
    """A parameter container; it has no forward of its own."""

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return self._modules[str(i)]

    def __iter__(self):
        return iter(self._modules.values())


def _init_tensor(entry):
    """Initialisation in the spirit of the reference (trunc-normal/xavier weights, zero biases, ones for
    norms, 0.1 layer scale, focal-prior class bias); exact init streams are not part of the contract."""
    shape, role = entry.shape, entry.role
    if role in ("linear", "conv", "convT", "class", "query_feat", "bbox_last", "sampling_offsets", "attention_weights"):
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = torch.empty(shape)
        nn.init.trunc_normal_(t, std=min(0.02 * math.sqrt(768.0 / max(fan_in, 1)) + 0.01, 0.05))
        return t
    if role in ("norm_weight", "bn_weight", "bn_var"):
        return torch.ones(shape)
    if role == "layer_scale":
        return torch.full(shape, 0.1)
    if role == "pos_embed":
        t = torch.empty(shape)
        nn.init.trunc_normal_(t, std=0.02)
        return t
    if role == "class_bias":
        return torch.full(shape, -math.log((1 - 0.01) / 0.01))
    if role == "bn_count":
        return torch.zeros((), dtype=torch.int64)
    return torch.zeros(shape)


def attach_entries(root, entries, strip=""):
    for e in entries:
        name = e.name[len(strip):] if strip and e.name.startswith(strip) else e.name
        parts = name.split(".")
        node = root
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, Holder())
            node = node._modules[p]
        t = _init_tensor(e)
        if e.kind == "param":
            node.register_parameter(parts[-1], nn.Parameter(t))
        else:
            node.register_buffer(parts[-1], t)
