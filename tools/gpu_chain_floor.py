"""Per-launch cost of a chain of dependent small conv launches replayed from a CUDA graph (the floor of the one-image
generator forward): python tools/gpu_chain_floor.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import ops  # noqa: E402


def chain(what, fn, n=40):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n):
                fn()
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    print("%-46s %6.2f us per launch (graph of %d dependent launches)" % (what, e0.elapsed_time(e1) * 1e3 / (10 * n), n))


def main():
    gen = torch.Generator().manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(s, generator=gen) * sc).to("cuda", torch.float16)
    a, w = r(1, 8, 8, 32), r(128, 9 * 32, sc=0.1)
    out = torch.empty((1, 8, 8, 128), device="cuda", dtype=torch.float16)
    chain("mlp_shared 8x8 (1 CTA, resident weights)",
          lambda: ops.conv_igemm([dict(a=a, w=w, C=32, taps=9)], 128, 128, act=ops.ACT_RELU, out=out))
    a2, w2 = r(1, 8, 8, 1024), r(1024, 9 * 1024, sc=0.01)
    o2 = torch.empty((1, 8, 8, 1024), device="cuda", dtype=torch.float32)
    chain("conv 1024->1024 8x8, BN=64, 16 CTAs",
          lambda: ops.conv_igemm([dict(a=a2, w=w2, C=1024, taps=9)], 1024, 64, out=o2))
    chain("the same, ksplit=8 (128 CTAs)",
          lambda: ops.conv_igemm([dict(a=a2, w=w2, C=1024, taps=9)], 1024, 64, out=o2, ksplit=8))
    chain("the same, ksplit=4 (64 CTAs)",
          lambda: ops.conv_igemm([dict(a=a2, w=w2, C=1024, taps=9)], 1024, 64, out=o2, ksplit=4))
    a3, w3 = r(1, 64, 64, 256), r(256, 9 * 256, sc=0.02)
    o3 = torch.empty((1, 64, 64, 256), device="cuda", dtype=torch.float32)
    chain("conv 256->256 64x64, BN=128, 64 CTAs",
          lambda: ops.conv_igemm([dict(a=a3, w=w3, C=256, taps=9)], 256, 128, out=o3))
    chain("the same, ksplit=2 (128 CTAs)",
          lambda: ops.conv_igemm([dict(a=a3, w=w3, C=256, taps=9)], 256, 128, out=o3, ksplit=2))
    x = torch.zeros((1 << 10,), device="cuda")
    chain("torch x.add_(1) on 1 K floats (reference floor)", lambda: x.add_(1.0))


if __name__ == "__main__":
    main()
