"""CPU baselines of the secondary paths, timed on the oracle ports (test infrastructure; run by hand next to
tools/bench_paths.py, not collected by pytest):   python tests/_cpu_baselines.py --what blend [--n 2]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ctrlhair_b200 import synth  # noqa: E402
from oracle import blend_oracle as bo  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="blend")
    ap.add_argument("--n", type=int, default=2)
    a = ap.parse_args()
    if "blend" in a.what:
        cases = [synth.make_blend_case(256, 256, 900 + i) for i in range(a.n)]
        t0 = time.time()
        for face, gen, fp, tp in cases:
            res = gen.transpose(2, 0, 1).astype("float32") / 127.5 - 1
            bo.postprocess_blending(face, res, fp, tp)
        dt = (time.time() - t0) / a.n
        print(json.dumps({"path": "8f.2 postprocess_blending 256x256, oracle port (vectorised assembly + scipy spsolve)",
                          "images_per_s": 1.0 / dt, "cores": 1, "sample": "%d images" % a.n}))


if __name__ == "__main__":
    main()
