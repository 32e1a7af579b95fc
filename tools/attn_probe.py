"""Runs the ViT attention core of one config a few times through the C ABI (for ncu / quick timing).
    python tools/attn_probe.py --config medium --kind global --iters 3
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lw-detr_b200"))

from b200 import capi  # noqa: E402
from b200.config import CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--kind", default="global", choices=["global", "window"])
    ap.add_argument("--dtype", default="")
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    B = a.batch or {"tiny": 32, "small": 32, "medium": 64, "large": 32, "xlarge": 16}[a.config]
    dt = {"": torch.bfloat16 if a.config == "medium" else torch.float16, "fp16": torch.float16, "bf16": torch.bfloat16}[a.dtype]
    C, heads, T = cfg.vit_dim, cfg.vit_heads, cfg.tokens
    dh = C // heads
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * T, 3 * C, device="cuda", generator=g).to(dt)
    att = torch.empty(B * T, C, device="cuda", dtype=dt)
    nseq, seqlen = (B, T) if a.kind == "global" else (16 * B, T // 16)
    ts = []
    for _ in range(a.iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        capi.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], att, nseq, seqlen, heads, dh, dh ** -0.5)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(a.config, a.kind, "B", B, "dh", dh, str(dt), "us:", ["%.1f" % t for t in ts])


if __name__ == "__main__":
    main()
