"""Quick device-resident timing of the generator forward (development aid; bench.py is the contract).

python tools/gpu_quick_bench.py [--B 64] [--crop 256] [--steps 5] [--warmup 2]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ctrlhair_b200.generator import SeanGeneratorB200  # noqa: E402
from ctrlhair_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=64)
    ap.add_argument("--crop", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--table", action="store_true")
    a = ap.parse_args()
    sd = synth.make_state_dict()
    gen = SeanGeneratorB200(crop=a.crop, max_batch=a.B)
    gen.load_state_dict(sd)
    del sd
    labels = synth.make_labels(a.B, a.crop, "blocky").cuda()
    codes = synth.make_codes(a.B).cuda()
    out = torch.empty((a.B, 3, a.crop, a.crop), device="cuda")
    for _ in range(a.warmup):
        gen.forward_labels(labels, codes, seed=1, out=out)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    ev[0].record()
    for i in range(a.steps):
        gen.forward_labels(labels, codes, seed=2 + i, out=out)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.steps)]
    fl = gen.flops(a.B)
    best = min(ms)
    print("B=%d crop=%d  ms/step: %s" % (a.B, a.crop, ", ".join("%.2f" % m for m in ms)))
    print("best %.2f ms -> %.1f img/s, %.1f TFLOP/s issued (%.2f GFLOP/img), finite=%s" %
          (best, a.B / best * 1e3, fl / best / 1e9, fl / a.B / 1e9, bool(torch.isfinite(out).all())))
    if a.table:
        _, ms1, fl1 = gen.forward_timed(labels, codes, seed=3, out=out)
        names = gen.step_names(a.B)
        print("per-launch (CUDA events between launches): total %.2f ms" % sum(ms1))
        for n, m, f in zip(names, ms1, fl1):
            print("  %-34s %8.3f ms %8.1f TFLOP/s" % (n, m, f / (m * 1e-3) / 1e12 if m > 0 else 0.0))


if __name__ == "__main__":
    main()
