"""One style-encoder forward inside a cudaProfilerStart/Stop window (for an ncu launch list):
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
    --log-file out.csv python tools/gpu_zenc_table.py [B];  python tools/launch_table.py out.csv"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.zencoder import ZencoderB200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    z = ZencoderB200(max_batch=B).load_state_dict(synth.make_state_dict())
    img, lab = synth.make_image(B, 256).cuda(), synth.make_labels(B, 256, "blocky").cuda()
    for _ in range(2):
        z(img, lab)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    z(img, lab)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
