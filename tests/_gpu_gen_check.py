"""GPU parity probe of the whole generator against the CPU oracle (and the golden fixtures), with per-block taps.

python tests/_gpu_gen_check.py [crop] [B] [kind]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ctrlhair_b200 import _lib  # noqa: E402
from ctrlhair_b200.generator import SeanGeneratorB200  # noqa: E402
from oracle import sean_oracle as so  # noqa: E402
from ctrlhair_b200 import synth  # noqa: E402


def main():
    crop = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    kind = sys.argv[3] if len(sys.argv) > 3 else "blocky"
    t0 = time.time()
    sd = synth.make_state_dict()
    print("state dict %.1fs" % (time.time() - t0), flush=True)
    labels = synth.make_labels(B, crop, kind)
    codes = synth.make_codes(B)
    noise = synth.make_noise(B, crop)
    t0 = time.time()
    taps = {}
    ref = so.generator_forward(sd, labels, codes, noise, taps=taps)
    print("oracle %.1fs" % (time.time() - t0), flush=True)
    gpath = os.path.join(ROOT, "tests", "golden", "gen_c%d_b%d_%s.npz" % (crop, B, kind))
    if os.path.exists(gpath):
        g = np.load(gpath)
        print("oracle vs golden(reference): %.3e" % float((ref - torch.from_numpy(g["out"])).abs().max()))
    gen = SeanGeneratorB200(crop=crop, max_batch=B)
    t0 = time.time()
    gen.load_state_dict(sd)
    torch.cuda.synchronize()
    print("pack+bind %.1fs, blob %.1f MB, launches %d, flops/img %.2f G" %
          (time.time() - t0, gen.blob_bytes() / 1e6, gen.launches(), gen.flops(B) / B / 1e9), flush=True)
    lab_d, codes_d, nz_d = labels.cuda(), codes.cuda(), synth.flatten_noise(noise).cuda()
    names = ["x_fc"] + ["x_" + b[0] for b in synth.BLOCKS]
    outs = {}
    for impl, iname in ((_lib.IMPL_SIMT_DEBUG, "simt"), (_lib.IMPL_TCGEN05, "tcgen05")):
        gen.impl = impl
        t0 = time.time()
        out = gen.forward_labels(lab_d, codes_d, noise=nz_d)
        torch.cuda.synchronize()
        dt = time.time() - t0
        outs[iname] = out.cpu()
        e = outs[iname] - ref
        print("%s: %.3fs  max|d|/max|ref| = %.3e  rel-L2 = %.3e  nonfinite=%d" %
              (iname, dt, float(e.abs().max() / ref.abs().max()), float(e.norm() / ref.norm()),
               int((~torch.isfinite(outs[iname])).sum())), flush=True)
        for n in names:
            t = taps[n]
            Bc, C, r, _ = t.shape
            got = gen.debug_tensor(n, (Bc, r, r, C), B).float().cpu().permute(0, 3, 1, 2)
            want = t
            if n == "x_up_3":
                want = torch.nn.functional.leaky_relu(t, 0.2)
            e = got - want
            print("   %-14s max|d|/max = %.3e  rel-L2 = %.3e" %
                  (n, float(e.abs().max() / want.abs().max()), float(e.norm() / want.norm())), flush=True)
    print("simt vs tcgen05: %.3e" % float((outs["simt"] - outs["tcgen05"]).abs().max()))
    # host-buffer entry point and device-drawn noise
    gen.impl = _lib.IMPL_TCGEN05
    oh = gen.forward_host(labels, codes, noise=synth.flatten_noise(noise))
    print("forward_host vs forward_labels: %.3e" % float((oh - outs["tcgen05"]).abs().max()))
    o2 = gen.forward_labels(lab_d, codes_d, noise=None, seed=7).cpu()
    print("device noise: finite=%s  max|d vs injected| = %.3f" %
          (bool(torch.isfinite(o2).all()), float((o2 - outs["tcgen05"]).abs().max())))


if __name__ == "__main__":
    main()
