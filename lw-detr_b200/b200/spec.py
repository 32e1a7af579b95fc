"""Checkpoint format of LW-DETR = the state_dict names and shapes of models.lwdetr.LWDETR.

`param_spec(cfg)` enumerates every entry (parameters and BatchNorm buffers) the reference model
registers for a configuration (reference: models/lwdetr.py:38-101, models/transformer.py:128-189,
430-464, models/backbone/vit.py:225-326, models/backbone/projector.py:141-212,
models/ops/modules/ms_deform_attn.py:37-77).  The drop-in module, the weight packer and the synthetic
weight generator are all driven by this one table; tests/test_spec.py pins it against the real
reference's state_dict().
"""
from collections import namedtuple

Entry = namedtuple("Entry", "name shape kind role")   # kind: 'param' | 'buffer' | 'buffer_i64'


def _linear(out, prefix, n_out, n_in, role="linear"):
    out.append(Entry(prefix + ".weight", (n_out, n_in), "param", role))
    out.append(Entry(prefix + ".bias", (n_out,), "param", role + "_bias"))


def _norm(out, prefix, n):
    out.append(Entry(prefix + ".weight", (n,), "param", "norm_weight"))
    out.append(Entry(prefix + ".bias", (n,), "param", "norm_bias"))


def _convx(out, prefix, c_out, c_in, k):
    out.append(Entry(prefix + ".conv.weight", (c_out, c_in, k, k), "param", "conv"))
    out.append(Entry(prefix + ".bn.weight", (c_out,), "param", "bn_weight"))
    out.append(Entry(prefix + ".bn.bias", (c_out,), "param", "bn_bias"))
    out.append(Entry(prefix + ".bn.running_mean", (c_out,), "buffer", "bn_mean"))
    out.append(Entry(prefix + ".bn.running_var", (c_out,), "buffer", "bn_var"))
    out.append(Entry(prefix + ".bn.num_batches_tracked", (), "buffer_i64", "bn_count"))


def _mlp(out, prefix, dims, last_role="linear"):
    for i in range(len(dims) - 1):
        role = last_role if i == len(dims) - 2 else "linear"
        _linear(out, "%s.layers.%d" % (prefix, i), dims[i + 1], dims[i], role)


def sampling_out_dim(cfg, scale):
    """Channels each tap contributes to a projector level (projector.py:165-200)."""
    C = cfg.vit_dim
    if scale == "P3":
        return C // 4 if C > 512 else C // 2
    return C


def param_spec(cfg):
    out = []
    d, C, nq, G = cfg.hidden_dim, cfg.vit_dim, cfg.num_queries, cfg.group_detr
    L, P, M = cfg.n_levels, cfg.dec_n_points, cfg.ca_nheads
    # ---- transformer (decoder + two-stage heads)
    for i in range(cfg.dec_layers):
        p = "transformer.decoder.layers.%d" % i
        out.append(Entry(p + ".self_attn.in_proj_weight", (3 * d, d), "param", "linear"))
        out.append(Entry(p + ".self_attn.in_proj_bias", (3 * d,), "param", "linear_bias"))
        _linear(out, p + ".self_attn.out_proj", d, d)
        _norm(out, p + ".norm1", d)
        _linear(out, p + ".cross_attn.sampling_offsets", M * L * P * 2, d, "sampling_offsets")
        _linear(out, p + ".cross_attn.attention_weights", M * L * P, d, "attention_weights")
        _linear(out, p + ".cross_attn.value_proj", d, d)
        _linear(out, p + ".cross_attn.output_proj", d, d)
        _linear(out, p + ".linear1", cfg.dim_feedforward, d)
        _linear(out, p + ".linear2", d, cfg.dim_feedforward)
        _norm(out, p + ".norm2", d)
        _norm(out, p + ".norm3", d)
    _norm(out, "transformer.decoder.norm", d)
    _mlp(out, "transformer.decoder.ref_point_head", (2 * d, d, d))
    for g in range(G):
        _linear(out, "transformer.enc_output.%d" % g, d, d)
    for g in range(G):
        _norm(out, "transformer.enc_output_norm.%d" % g, d)
    for g in range(G):
        _mlp(out, "transformer.enc_out_bbox_embed.%d" % g, (d, d, d, 4), "bbox_last")
    for g in range(G):
        _linear(out, "transformer.enc_out_class_embed.%d" % g, cfg.num_classes, d, "class")
    # ---- detection heads / queries
    _linear(out, "class_embed", cfg.num_classes, d, "class")
    _mlp(out, "bbox_embed", (d, d, d, 4), "bbox_last")
    out.append(Entry("refpoint_embed.weight", (nq * G, 4), "param", "refpoint"))
    out.append(Entry("query_feat.weight", (nq * G, d), "param", "query_feat"))
    # ---- ViT encoder
    e = "backbone.0.encoder"
    out.append(Entry(e + ".pos_embed", (1, 197, C), "param", "pos_embed"))
    out.append(Entry(e + ".patch_embed.proj.weight", (C, 3, cfg.patch, cfg.patch), "param", "conv"))
    out.append(Entry(e + ".patch_embed.proj.bias", (C,), "param", "linear_bias"))
    for i in range(cfg.vit_depth):
        b = "%s.blocks.%d" % (e, i)
        out.append(Entry(b + ".gamma_1", (C,), "param", "layer_scale"))
        out.append(Entry(b + ".gamma_2", (C,), "param", "layer_scale"))
        _norm(out, b + ".norm1", C)
        out.append(Entry(b + ".attn.q_bias", (C,), "param", "linear_bias"))
        out.append(Entry(b + ".attn.v_bias", (C,), "param", "linear_bias"))
        out.append(Entry(b + ".attn.qkv.weight", (3 * C, C), "param", "linear"))
        _linear(out, b + ".attn.proj", C, C)
        _norm(out, b + ".norm2", C)
        _linear(out, b + ".mlp.fc1", 4 * C, C)
        _linear(out, b + ".mlp.fc2", C, 4 * C)
    # ---- multi-level projector
    pr = "backbone.0.projector"
    ntap = len(cfg.taps)
    for lvl, scale in enumerate(cfg.projector_scale):
        for t in range(ntap):
            s = "%s.stages_sampling.%d.%d" % (pr, lvl, t)
            if scale == "P3":
                if C > 512:
                    _convx(out, s + ".0", C // 2, C, 1)
                    out.append(Entry(s + ".1.weight", (C // 2, C // 4, 2, 2), "param", "convT"))
                    out.append(Entry(s + ".1.bias", (C // 4,), "param", "linear_bias"))
                else:
                    out.append(Entry(s + ".0.weight", (C, C // 2, 2, 2), "param", "convT"))
                    out.append(Entry(s + ".0.bias", (C // 2,), "param", "linear_bias"))
            elif scale == "P5":
                _convx(out, s + ".0", C, C, 3)
    for lvl, scale in enumerate(cfg.projector_scale):
        st = "%s.stages.%d" % (pr, lvl)
        c = d // 2
        _convx(out, st + ".0.cv1", 2 * c, sampling_out_dim(cfg, scale) * ntap, 1)
        _convx(out, st + ".0.cv2", d, 5 * c, 1)
        for j in range(3):
            _convx(out, "%s.0.m.%d.cv1" % (st, j), c, c, 3)
            _convx(out, "%s.0.m.%d.cv2" % (st, j), c, c, 3)
        _norm(out, st + ".1", d)
    return out


def num_parameters(cfg):
    n = 0
    for e in param_spec(cfg):
        if e.kind == "param":
            k = 1
            for s in e.shape:
                k *= s
            n += k
    return n
