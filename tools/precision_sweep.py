"""GPU sweep over the precision policies of SeanGeneratorB200: step time at B = 64, 256x256 and the error against the
oracle on the test-suite cases.  python tools/precision_sweep.py [policy ...]  -> one JSON line per policy."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.generator import PRECISION_POLICIES, SeanGeneratorB200  # noqa: E402
from oracle import sean_oracle as so  # noqa: E402  (development tool: the oracle is the checker here)


def errs(got, ref):
    d = got - ref
    return [float(d.abs().max() / ref.abs().max()), float(d.norm() / ref.norm())]


def main():
    policies = sys.argv[1:] or ["fast", "shortcut", "h1", "parity", "full", "margin"]
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_state_dict()
    B = 64
    L, Cd, N = synth.make_labels(B, 256, "blocky"), synth.make_codes(B), synth.make_noise(B, 256)
    picks = (0, 31, 32, 63)
    refs = {i: so.generator_forward(sd, L[i:i + 1], Cd[i:i + 1], [p[i:i + 1] for p in N]) for i in picks}
    g = np.load(os.path.join(ROOT, "tests", "golden", "gen_c256_b1_blocky_ui.npz"))
    small = {}
    for kind in ("blocky", "iid"):
        l, c, n = synth.make_labels(2, 64, kind), synth.make_codes(2), synth.make_noise(2, 64)
        small[kind] = (l, c, n, so.generator_forward(sd, l, c, n))
    l5, c5, n5 = synth.make_labels(1, 512, "blocky", seed=21), synth.make_codes(1, seed=22), synth.make_noise(1, 512)
    r5 = so.generator_forward(sd, l5, c5, n5)
    for pol in policies:
        flags = PRECISION_POLICIES[pol] if pol in PRECISION_POLICIES else int(pol, 0)
        gen = SeanGeneratorB200(crop=256, max_batch=B, precision=flags).load_state_dict(sd)
        out = gen.forward_labels(L.cuda(), Cd.cuda(), noise=synth.flatten_noise(N).cuda())
        res = {"policy": pol, "flags": flags, "b64": {i: errs(out[i:i + 1].cpu(), refs[i]) for i in picks}}
        ui = gen.forward_labels(torch.from_numpy(g["labels"]).cuda(), synth.make_codes(1).cuda(),
                                noise=synth.flatten_noise(synth.make_noise(1, 256)).cuda()).cpu()
        res["c256_ui_golden"] = errs(ui, torch.from_numpy(g["out"]))
        lab, cod = L.cuda(), Cd.cuda()
        for i in range(3):
            gen.forward_labels(lab, cod, seed=i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(8):
            gen.forward_labels(lab, cod, seed=10 + i)
        e1.record()
        torch.cuda.synchronize()
        res["ms_per_step"] = e0.elapsed_time(e1) / 8
        res["img_per_s"] = B * 8 / (e0.elapsed_time(e1) * 1e-3)
        _, ms, fl = gen.forward_timed(lab, cod, seed=3)
        names = gen.step_names(B)
        res["launch_ms"] = {n: round(m, 4) for n, m in zip(names, ms) if m > 0.25}
        res["conv_ms"] = sum(ms)
        del gen
        torch.cuda.empty_cache()
        g64 = SeanGeneratorB200(crop=64, max_batch=2, precision=flags).load_state_dict(sd)
        for kind, (l, c, n, r) in small.items():
            res["c64_" + kind] = errs(g64.forward_labels(l.cuda(), c.cuda(), noise=synth.flatten_noise(n).cuda()).cpu(), r)
        del g64
        g512 = SeanGeneratorB200(crop=512, max_batch=1, precision=flags).load_state_dict(sd)
        res["c512"] = errs(g512.forward_labels(l5.cuda(), c5.cuda(), noise=synth.flatten_noise(n5).cuda()).cpu(), r5)
        del g512
        torch.cuda.empty_cache()
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
