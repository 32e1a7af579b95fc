"""Run a few eager (non-graph) forwards of one configuration - the target command for ncu captures.
    ncu ... python tools/run_forward.py --config small --batch 32 --dtype fp16 --iters 3
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
from b200.synth import synth_images, synth_state_dict  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="small")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--dtype", default="fp16")
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
cfg = CONFIGS[a.config]
eng = capi.Engine(cfg, {"fp16": torch.float16, "bf16": torch.bfloat16}[a.dtype])
eng.load_state_dict(synth_state_dict(cfg, 1))
x = synth_images(a.batch, 0).cuda()
for _ in range(a.iters):
    out = eng.forward(x, want_aux=False)
torch.cuda.synchronize()
print("ops per forward:", len(eng.ops()), "finite:", bool(torch.isfinite(out["pred_logits"]).all()))
