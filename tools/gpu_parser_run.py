"""Runs the GPU face parser a few times (for ncu launch lists / timing): python tools/gpu_parser_run.py [B] [reps]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.bisenet import BiSeNetB200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    net = BiSeNetB200(max_batch=B).load_state_dict(synth.make_bisenet_state_dict())
    rng = np.random.default_rng(0)
    img = torch.from_numpy(rng.integers(0, 256, (B, 256, 256, 3), dtype=np.uint8)).cuda()
    for _ in range(2):
        net.get_mask_device(img, 256)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.get_mask_device(img, 256)
    e1.record()
    torch.cuda.synchronize()
    print("face parser B=%d: %.3f ms per batch (resize + network + tail), %.0f img/s" %
          (B, e0.elapsed_time(e1) / reps, B * reps / (e0.elapsed_time(e1) * 1e-3)))


if __name__ == "__main__":
    main()
