"""Poisson kernels under compute-sanitizer racecheck at a reduced iteration cap (the 650-iteration solves do not finish
under the tool in reasonable time; the shared-memory / DSMEM protocol per iteration is the same):
    compute-sanitizer --tool racecheck python tools/gpu_racecheck_poisson.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import blend, synth  # noqa: E402


def main():
    cap = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    for (h, w) in ((256, 256), (96, 80)):       # cluster-resident kernel (256 wide) and the first-generation kernel
        face, gen, fp, tp = synth.make_blend_case(h, w, 900)
        mask = 1 - blend.blend_mask(torch.from_numpy(tp).cuda(), torch.from_numpy(fp).cuda())
        out, stats = blend.poisson_blending(face, gen, mask, max_iter=cap, return_stats=True)
        torch.cuda.synchronize()
        print("poisson %dx%d: %d iterations run, output finite %s" % (h, w, int(stats[..., 0].max()),
                                                                      bool(np.isfinite(out.cpu().numpy()).all())))


if __name__ == "__main__":
    main()
