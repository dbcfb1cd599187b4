"""Per-launch times of the generator schedule at a small batch (default B = 1): python tools/gpu_b1_table.py [B] [policy]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.generator import SeanGeneratorB200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    pol = sys.argv[2] if len(sys.argv) > 2 else "parity"
    gen = SeanGeneratorB200(crop=256, max_batch=B, precision=pol).load_state_dict(synth.make_state_dict())
    lab, cod = synth.make_labels(B, 256, "blocky").cuda(), synth.make_codes(B).cuda()
    for i in range(3):
        gen.forward_labels(lab, cod, seed=i)
    best = None
    for rep in range(5):
        _, ms, fl = gen.forward_timed(lab, cod, seed=3)
        if best is None or sum(ms) < sum(best):
            best = ms
    names = gen.step_names(B)
    print("B=%d policy=%s: %d launches, %.3f ms between-launch events" % (B, pol, len(best), sum(best)))
    for n, m, f in zip(names, best, fl):
        print("  %-36s %8.1f us  %7.1f TFLOP/s" % (n, m * 1e3, f / (m * 1e-3) / 1e12 if m > 0 else 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for mode in (False, True):
        for i in range(5):
            gen.forward_labels(lab, cod, seed=i, graph=mode)
        torch.cuda.synchronize()
        e0.record()
        for i in range(50):
            gen.forward_labels(lab, cod, seed=i, graph=mode)
        e1.record()
        torch.cuda.synchronize()
        print("graph=%s: %.3f ms per forward" % (mode, e0.elapsed_time(e1) / 50))


if __name__ == "__main__":
    main()
