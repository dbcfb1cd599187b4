"""One encode + one decode of the shape nets inside a cudaProfilerStart/Stop window (for an ncu launch list):
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv \
    --log-file out.csv python tools/gpu_shape_table.py [B];  python tools/launch_table.py out.csv"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.shape import ShapeGeneratorB200  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    s = ShapeGeneratorB200(max_batch=B).load_state_dict(synth.make_shape_state_dict())
    hair, face = synth.make_shape_inputs(B)
    hair, face = hair.cuda(), face.cuda()
    for _ in range(2):
        hc, fc = s.forward_hair_encoder(hair, testing=True), s.forward_face_encoder(face)
        s.forward_decode_by_code(hc, fc)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    hc, fc = s.forward_hair_encoder(hair, testing=True), s.forward_face_encoder(face)
    torch.cuda.synchronize()
    s.forward_decode_by_code(hc, fc)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
