"""One Poisson solve at batch B inside a cudaProfilerStart/Stop window (for ncu): python tools/gpu_poisson_once.py [B]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import blend, synth  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    cases = [synth.make_blend_case(256, 256, 900 + i) for i in range(B)]
    face = torch.from_numpy(np.stack([c[0] for c in cases])).cuda()
    gen = torch.from_numpy(np.stack([c[1] for c in cases])).cuda()
    mask = 1 - blend.blend_mask(torch.from_numpy(np.stack([c[3] for c in cases])),
                                torch.from_numpy(np.stack([c[2] for c in cases])))
    blend.poisson_blending(face, gen, mask)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    _, st = blend.poisson_blending(face, gen, mask, return_stats=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("iterations", st[..., 0].flatten().tolist())


if __name__ == "__main__":
    main()
