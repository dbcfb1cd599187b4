"""Timing staircase of the Poisson solver over the batch size (how many 8-CTA clusters run at once, time per iteration).
python tools/poisson_sweep.py            (CHB_POISSON_V1=1 selects the first-generation kernel)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import blend, synth  # noqa: E402


def main():
    cases = [synth.make_blend_case(256, 256, 900 + i) for i in range(32)]
    face = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    gen = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    mask = 1 - blend.blend_mask(torch.from_numpy(np.stack([c[3] for c in cases])),
                                torch.from_numpy(np.stack([c[2] for c in cases])))
    for B in (1, 2, 3, 4, 5, 6, 7, 8, 11, 12, 16, 22, 32):
        f, g, m = face[:B].contiguous(), gen[:B].contiguous(), mask[:B].contiguous()
        _, st = blend.poisson_blending(f, g, m, return_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            blend.poisson_blending(f, g, m)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        it = float(st[..., 0].mean())
        print("B=%2d systems=%2d  %.3f ms  iterations %.0f  -> %.2f us per iteration if all systems run at once" %
              (B, 3 * B, ms, it, ms * 1e3 / it))


if __name__ == "__main__":
    main()
