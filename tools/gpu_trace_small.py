"""Where the time of a small conv launch goes: SM-clock timestamps of CTA 0 (tuning build with -DCHB_TRACE).
  VARIANT_FLAGS=-DCHB_TRACE python tools/build_variant.py trace
  CHB_LIB_PATH=variants/libtrace.so python tools/gpu_trace_small.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import _lib, ops  # noqa: E402

NAMES = ["entry", "prologue done", "producer: first tile", "mma: accumulator free", "mma: tile committed",
         "epilogue: accumulator full", "epilogue: tile stored", "all roles done", "tmem freed", "mma: resident weights in"]


def trace():
    buf = (C.c_ulonglong * 16)()
    lib = _lib.load()
    torch.cuda.synchronize()
    assert lib.chb_debug_trace_read(buf) == 0
    return list(buf)


def show(what, fn, reps=3):
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = trace()
    print("%s: %.1f us per launch back to back" % (what, e0.elapsed_time(e1) * 1e3 / 20))
    order = sorted((v, k) for k, v in enumerate(t[:10]) if v)
    for v, k in order:
        print("   +%7d cycles  %s" % (v - t[0], NAMES[k]))
    cta = (C.c_ulonglong * 480)()
    assert _lib.load().chb_debug_trace_cta_read(cta) == 0
    rows = [(cta[3 * i], cta[3 * i + 1], cta[3 * i + 2]) for i in range(160) if cta[3 * i + 2] > cta[3 * i] > 0]
    t_end = max(r[2] for r in rows)
    rows = [r for r in rows if r[2] > t_end - 200000]          # the CTAs of the last launch
    t0 = min(r[0] for r in rows)
    print("   last launch: %d CTAs, starts spread over %.1f us, first start -> last exit %.1f us; per CTA: entry -> "
          "accumulator full %.1f us (max), -> exit %.1f us (max)" %
          (len(rows), (max(r[0] for r in rows) - t0) / 1e3, (t_end - t0) / 1e3,
           max(r[1] - r[0] for r in rows) / 1e3, max(r[2] - r[0] for r in rows) / 1e3))
    ks = (C.c_ulonglong * 128)()
    _lib.load().chb_debug_trace_ks_read.argtypes = [C.c_void_p]
    assert _lib.load().chb_debug_trace_ks_read(ks) == 0
    pts = ["enter", "partial stored", "fence", "atomic", "bar 2", "fence 2", "summed + st issued", "st waited"]
    for sp in range(16):
        v = [ks[8 * sp + k] for k in range(8)]
        if v[0] > t_end - 200000:
            print("   tile 0 split %2d: " % sp + "  ".join("%s +%.1f" % (pts[k], (v[k] - t0) / 1e3) for k in range(8)
                                                           if v[k] >= v[0]))
    late = sorted(rows, key=lambda r: r[2])[-3:]
    print("   the three CTAs that exit last: " + ", ".join("start +%.1f full +%.1f exit +%.1f" %
          ((r[0] - t0) / 1e3, (r[1] - t0) / 1e3, (r[2] - t0) / 1e3) for r in late))


def main():
    _lib.load().chb_debug_trace_read.argtypes = [C.c_void_p]
    _lib.load().chb_debug_trace_cta_read.argtypes = [C.c_void_p]
    gen = torch.Generator().manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(s, generator=gen) * sc).to("cuda", torch.float16)
    # a tiny weight-stationary launch (mlp_shared at 8x8: one-hot 32 -> 128), one CTA
    a, w = r(1, 8, 8, 32), r(128, 9 * 32, sc=0.1)
    show("mlp_shared 8x8, 1 CTA", lambda: ops.conv_igemm([dict(a=a, w=w, C=32, taps=9)], 128, 128, act=ops.ACT_RELU))
    # a 1024 -> 1024 conv at 8x8, N tile 64 (16 CTAs), unsplit and split 8 ways
    a2, w2 = r(1, 8, 8, 1024), r(1024, 9 * 1024, sc=0.01)
    show("conv 1024->1024 8x8 BN=64", lambda: ops.conv_igemm([dict(a=a2, w=w2, C=1024, taps=9)], 1024, 64,
                                                              out_dtype=torch.float32))
    show("the same, ksplit=8", lambda: ops.conv_igemm([dict(a=a2, w=w2, C=1024, taps=9)], 1024, 64,
                                                      out_dtype=torch.float32, ksplit=8))
    show("the same, ksplit=4", lambda: ops.conv_igemm([dict(a=a2, w=w2, C=1024, taps=9)], 1024, 64,
                                                      out_dtype=torch.float32, ksplit=4))
    a3, w3 = r(1, 16, 16, 128), r(256, 9 * 128, sc=0.03)
    show("conv 128->256 16x16 BN=256, 2 CTAs", lambda: ops.conv_igemm([dict(a=a3, w=w3, C=128, taps=9)], 256, 256,
                                                                      out_dtype=torch.float32))


if __name__ == "__main__":
    main()
