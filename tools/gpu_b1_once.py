"""One generator forward at a small batch inside a cudaProfilerStart/Stop window (for an ncu launch list):
ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file out.csv python tools/gpu_b1_once.py [B]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.generator import SeanGeneratorB200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    gen = SeanGeneratorB200(crop=256, max_batch=B).load_state_dict(synth.make_state_dict())
    lab, cod = synth.make_labels(B, 256, "blocky").cuda(), synth.make_codes(B).cuda()
    for i in range(3):
        gen.forward_labels(lab, cod, seed=i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    gen.forward_labels(lab, cod, seed=7)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("\n".join(gen.step_names(B)))


if __name__ == "__main__":
    main()
