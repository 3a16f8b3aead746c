"""BASELINE configs[1]: SigLIP-so400m tower only (26 blocks + final LayerNorm, no head), batch 1 -> 512, 384x384 inputs.
Through the VisRAG `vpm` (dynamic_img_pad) a 384 px image pads to 392 -> 28x28 = 784 tokens; here the
synthetic input is generated at 392x392 directly. Prints images/s and model TFLOP/s per batch size.
  python tools/bench_vit_sweep.py [--max-batch 512]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from visrag_b200.config import VisRAGConfig  # noqa: E402
from visrag_b200.encoder import VisRAGEngine  # noqa: E402
from visrag_b200.weights import random_state_dict_device  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-batch", type=int, default=512)
    ap.add_argument("--px", type=int, default=392)
    a = ap.parse_args()
    cfg = VisRAGConfig.full()
    eng = VisRAGEngine(cfg, random_state_dict_device(cfg, 2024, "cuda:0"))
    n = (a.px // cfg.patch_size) ** 2
    flop_per_image = 793_046_016 * n + 119_808 * n * n  # F_vit(N)
    g = torch.Generator(device="cuda").manual_seed(0)
    bs = 1
    while bs <= a.max_batch:
        px = torch.randint(0, 256, (bs, a.px, a.px, 3), dtype=torch.uint8, device="cuda", generator=g)
        for _ in range(3):
            eng.vit_tokens(px)
        torch.cuda.synchronize()
        reps = max(3, min(50, 2048 // bs))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            eng.vit_tokens(px)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"batch {bs:4d} x {n} tokens: {ms:8.3f} ms  {bs / ms * 1e3:9.1f} images/s  {flop_per_image * bs / ms / 1e9:7.1f} TFLOP/s",
              flush=True)
        bs *= 2


if __name__ == "__main__":
    main()
