"""Where config 3's wall time goes (host clock with a sync after each stage): python tools/gpu_config3_breakdown.py"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.backend import BackendB200  # noqa: E402
import bench_paths as bp  # noqa: E402


def main():
    B = 32
    faces, _ = bp.load_example_faces(B)
    img_h = torch.from_numpy(faces).pin_memory()
    out_h = torch.empty((B, 256, 256, 3), dtype=torch.uint8).pin_memory()
    be = BackendB200(synth.make_state_dict(), synth.make_shape_state_dict(), synth.make_ct_state_dicts(),
                     median_codes=synth.make_codes(1, seed=4321)[0], max_batch=B, blending=True,
                     parsing_sd=synth.make_bisenet_state_dict())
    sync = torch.cuda.synchronize
    stages = {}

    def t(name, fn):
        sync(); t0 = time.perf_counter(); r = fn(); sync()
        stages.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
        return r
    for it in range(6):
        img_d = t("h2d", lambda: img_h.cuda(non_blocking=True))
        mask = t("get_mask (parser)", lambda: be.get_mask(img_d))
        t("set_input_img (encoders, colour, hsv)", lambda: be.set_input_img(img_d, mask))
        for blend in (False, True):
            be.blending = blend
            o = t("output, blending=%s" % blend, lambda: be.output())
        t("d2h", lambda: out_h.copy_(o, non_blocking=True))
    for k, v in stages.items():
        print("%-42s %7.2f ms (min of %d after warm-up)" % (k, min(v[2:]), len(v) - 2))
    print("unknown fraction of the blend masks: see chb_poisson stats; mask classes:", torch.unique(be.cur_mask).tolist()[:10])


if __name__ == "__main__":
    main()
