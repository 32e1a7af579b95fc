"""A few plain (no CUDA graph) forwards of one configuration - the target of `ncu -k regex:<kernel> -s N -c M` captures of
in-model launches (the launch order of a forward is Engine::ops(); `--list` prints it).
    python tools/one_forward.py --config small --batch 32 --dtype fp16 [--n 2] [--list]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lw-detr_b200"))

from b200.config import CONFIGS  # noqa: E402
from b200.synth import synth_images, synth_state_dict  # noqa: E402
from models.lwdetr import LWDETR  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="small")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--n", type=int, default=2)
    ap.add_argument("--list", action="store_true")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    dt = {"fp16": torch.float16, "bf16": torch.bfloat16}[a.dtype]
    model = LWDETR(cfg, compute_dtype=dt).eval()
    model.load_state_dict(synth_state_dict(cfg, 1), strict=True)
    model.to("cuda:0")
    eng = model.engine()
    eng.set_option("cuda_graph", 0)
    x = synth_images(a.batch, 0).to("cuda:0")
    for _ in range(a.n):
        eng.forward(x, want_aux=False)
    torch.cuda.synchronize()
    if a.list:
        for i, o in enumerate(eng.ops()):
            print(i, o)


if __name__ == "__main__":
    main()
