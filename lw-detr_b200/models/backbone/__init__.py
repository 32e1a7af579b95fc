"""`models.backbone` surface callers touch (reference: models/backbone/__init__.py:11-63):
Joiner (an nn.Sequential of [backbone, position embedding]) and build_backbone(args)."""
import torch
from torch import nn

from b200.config import config_from_args
from b200.spec import param_spec

from .._tree import attach_entries


class PositionEmbeddingPlaceholder(nn.Module):
    """Joiner[1].  The sine position encodings are computed by the reference but never consumed by the
    decoder (transformer.py:466-517 ignores `pos`), so nothing is computed here."""

    def forward(self, *args, **kwargs):
        return None


class Backbone(nn.Module):
    """Holds the ViT encoder + projector parameters under the reference's names (backbone.0.*).
    The math runs inside the CUDA engine owned by LWDETR; calling this module directly is not supported."""

    def __init__(self, cfg):
        super().__init__()
        self.name = cfg.encoder
        self.projector_scale = list(cfg.projector_scale)
        entries = [e for e in param_spec(cfg) if e.name.startswith("backbone.0.")]
        attach_entries(self, entries, strip="backbone.0.")
        for blk in self.encoder.blocks.children():
            blk.drop_path = nn.Identity()          # touched by LWDETR.update_drop_path (lwdetr.py:205-210)
        self._export = False

    def export(self):
        self._export = True

    def forward(self, *args, **kwargs):
        raise RuntimeError("lwdetr_b200: the backbone runs inside the fused CUDA engine; call the LWDETR module")

    def get_named_param_lr_pairs(self, args, prefix: str = "backbone.0"):
        """Layer-wise lr decay groups for the ViT (same contract as backbone.py:173-233): block i gets
        lr_encoder * decay^(L+1-(i+1)), patch/pos embeddings decay^(L+1); no weight decay on
        gamma/pos_embed/bias/norm parameters."""
        n_layers = args.vit_encoder_num_layers
        out = {}
        for n, p in self.named_parameters():
            full = prefix + "." + n
            if "backbone.0.encoder" not in full or not p.requires_grad:
                continue
            layer_id = n_layers + 1
            if ".pos_embed" in full or ".patch_embed" in full:
                layer_id = 0
            elif ".blocks." in full:
                layer_id = int(full.split(".blocks.")[1].split(".")[0]) + 1
            lr = args.lr_encoder * args.lr_vit_layer_decay ** (n_layers + 1 - layer_id) * args.lr_component_decay ** 2
            no_wd = any(s in full for s in ("gamma", "pos_embed", "rel_pos", "bias", "norm"))
            out[full] = {"params": p, "lr": lr, "weight_decay": 0.0 if no_wd else args.weight_decay}
        return out


class Joiner(nn.Sequential):
    def __init__(self, backbone, position_embedding):
        super().__init__(backbone, position_embedding)
        self._export = False

    def export(self):
        self._export = True

    def forward(self, *args, **kwargs):
        raise RuntimeError("lwdetr_b200: the backbone runs inside the fused CUDA engine; call the LWDETR module")


def build_backbone(args, cfg=None):
    cfg = cfg if cfg is not None else config_from_args(args)
    return Joiner(Backbone(cfg), PositionEmbeddingPlaceholder())
