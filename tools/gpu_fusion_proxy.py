"""Upper bound on what computing `actv` inside the gamma/beta kernels could buy ("stage 1"), measured
with the existing operator at up_3's geometry (B = 64, 256 x 256):

  now     mlp_shared (one-hot 32 -> 3 x 128 channels, relu, fp16 store) + ace_s + ace_0 (N = 256) + ace_1 (N = 128)
  proxy   the three gamma/beta launches alone, each with EXTRA K-segments on the one-hot map whose tensor-core work
          equals recomputing its actv tile on the (16+2) x (8+2) halo: two 128-row MMA passes x K = 288 x N = 128
          = 36 MMAs of N = 128, i.e. K + 288 for the N = 256 launches and K + 576 for the N = 128 launch.

The proxy is generous to the fusion: it still READS a stored actv (the fused kernel would not, but it would have to
round-trip its actv tile TMEM -> registers -> swizzled shared memory and lose the double-buffered accumulator).  If the
proxy is not faster than "now", the fused kernel cannot be.   python tools/gpu_fusion_proxy.py"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import ops  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    B, S = 64, 256
    gen = torch.Generator().manual_seed(0)
    r16 = lambda *s, sc=1.0: (torch.randn(s, generator=gen) * sc).to("cuda", torch.float16)
    lab = torch.randint(0, 19, (B, S, S), generator=gen)
    onehot = F.one_hot(lab, 32).to("cuda", torch.float16)
    actv = torch.relu(r16(B, S, S, 384))                       # the three ACEs' slices of one mlp_shared output
    w_mlp = r16(384, 9 * 32, sc=0.1)
    actv_out = torch.empty((B, S, S, 384), device="cuda", dtype=torch.float16)
    res = {}
    res["mlp_shared"] = timed(lambda: ops.conv_igemm([dict(a=onehot, w=w_mlp, C=32, taps=9)], 384, 128,
                                                     act=ops.ACT_RELU, out=actv_out))
    w_extra = r16(256, 9 * 32, sc=0.1)

    def ace(C, x_shift, slot, extra):
        N = 2 * C
        weff = r16(B, N, 9 * 32, sc=0.1)
        wgb = r16(N, 9 * 128, sc=0.03)
        bias = torch.randn((N,), generator=gen).cuda() * 0.3
        xh = S >> x_shift
        x = torch.randn((B, xh, xh, C), generator=gen).cuda()
        noise = torch.randn((B, S, S), generator=gen).cuda()
        chan = torch.stack([torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.3,
                            torch.randn(C, generator=gen) * 0.1]).cuda()
        out = torch.empty((B, S, S, C), device="cuda", dtype=torch.float16)
        segs = [dict(a=onehot, w=weff, C=32, taps=9), dict(a=actv, w=wgb, C=128, ch_off=128 * slot, taps=9)]
        segs += [dict(a=onehot, w=w_extra[:N].contiguous(), C=32, taps=9) for _ in range(extra)]
        return timed(lambda: ops.conv_igemm(segs, N, 256 if N >= 256 else N, bias=bias, epi=ops.EPI_MODULATE,
                                            act=ops.ACT_LRELU, x=x, x_shift=x_shift, noise=noise, chan=chan, out=out))
    for name, C, xs, slot, extra in (("ace_s", 128, 1, 0, 1), ("ace_0", 128, 1, 1, 1), ("ace_1", 64, 0, 2, 2)):
        res[name] = ace(C, xs, slot, 0)
        res[name + " + actv-recompute MMAs"] = ace(C, xs, slot, extra)
    for k, v in res.items():
        print("%-34s %.3f ms" % (k, v))
    now = res["mlp_shared"] + res["ace_s"] + res["ace_0"] + res["ace_1"]
    proxy = sum(v for k, v in res.items() if k.endswith("MMAs"))
    print("now (mlp_shared + three gamma/beta launches)      %.3f ms" % now)
    print("proxy for the fused kernels (upper bound on gain)  %.3f ms   -> %+.3f ms" % (proxy, proxy - now))


if __name__ == "__main__":
    main()
