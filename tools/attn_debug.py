"""Run the slot attention kernel on a few shapes with LWDETR_B200_DEBUG_WAIT=1 and print which barrier wait timed out."""
import os
import sys
os.environ.setdefault("LWDETR_B200_DEBUG_WAIT", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lw-detr_b200"))
import torch
from b200 import capi

shapes = [(8, 100, 4, 16), (2, 300, 4, 16), (2, 1600, 12, 16), (16, 100, 12, 32)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for nseq, seqlen, heads, dh in shapes:
    C = heads * dh
    g = torch.Generator(device="cuda").manual_seed(1)
    qkv = (torch.randn(nseq * seqlen, 3 * C, device="cuda", generator=g) * 1.5).half()
    out = torch.full((nseq * seqlen, C), float("nan"), device="cuda", dtype=torch.float16)
    try:
        capi.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], out, nseq, seqlen, heads, dh, dh ** -0.5)
        torch.cuda.synchronize()
        q, k, v = [t.float().reshape(nseq, seqlen, heads, dh).transpose(1, 2) for t in (qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:])]
        s = (q * dh ** -0.5) @ k.transpose(-2, -1)
        ref = (s.softmax(-1) @ v).transpose(1, 2).reshape(nseq * seqlen, C)
        err = (out.float() - ref).abs()
        print("shape", (nseq, seqlen, heads, dh), "max err %.3e" % err.max().item(), "nan rows", int(torch.isnan(out.float()).any(1).sum()),
              "rel l2 %.3e" % ((out.float() - ref).norm() / ref.norm()).item(), flush=True)
        if err.max().item() > 2e-2 or torch.isnan(out.float()).any():
            bad = (err > 2e-2) | torch.isnan(out.float())
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print("  bad rows (first 20):", rows[:20].tolist(), "... count", rows.numel(), " bad cols:", cols[:32].tolist(), flush=True)
    except Exception as e:
        print("shape", (nseq, seqlen, heads, dh), "FAILED:", e, flush=True)
        capi.lib().lwdetr_debug_dump()
        break
