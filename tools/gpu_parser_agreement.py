"""Label agreement of the GPU face parser with the fp32 oracle on other inputs than the golden one (the reference's
example faces, resized to 512x512 as my_parsing_util.py:35 does): python tools/gpu_parser_agreement.py [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.bisenet import BiSeNetB200  # noqa: E402
from oracle import bisenet_oracle as bno  # noqa: E402  (development tool: the oracle is the checker here)
import bench_paths as bp  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_bisenet_state_dict()
    net = BiSeNetB200(max_batch=1, swap_labels=False).load_state_dict(sd)
    faces, _ = bp.load_example_faces(n)
    rng = np.random.default_rng(5)
    inputs = [("imgs/*.png #%d" % i, net.resize_to_network(faces[i], 512)) for i in range(n)]
    inputs.append(("uniform noise", rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)))
    ag = []
    for name, im in inputs:
        img = np.asarray(im, dtype=np.uint8)[None]
        parsing, logits = net(torch.from_numpy(img).cuda(), return_logits=True)
        with torch.no_grad():
            ref_low = bno.bisenet_logits_lowres(sd, bno.normalise_image(img))
            ref_full = torch.nn.functional.interpolate(ref_low, (512, 512), mode="bilinear", align_corners=True)
        want = ref_full.argmax(1)[0].numpy()
        got = parsing[0].cpu().numpy()
        err = float((logits.cpu().permute(0, 3, 1, 2) - ref_low).abs().max())
        top2 = ref_full[0].topk(2, dim=0).values
        margin = (top2[0] - top2[1]).numpy()
        bad = got != want
        ag.append(float((got == want).mean()))
        print("%-18s agreement %.5f  logit max err %.2e of %.2f  largest margin among misses %.2e  classes %d" %
              (name, ag[-1], err, float(ref_low.abs().max()), float(margin[bad].max()) if bad.any() else 0.0,
               len(np.unique(want))), flush=True)
    print("min %.5f  mean %.5f" % (min(ag), float(np.mean(ag))))


if __name__ == "__main__":
    main()
