"""CPU study: which fp16 rounding sites of the packed schedule cost how much end-to-end error (development aid).

python tools/precision_study.py [--B 1] [--dist blocky]

Runs tests/_packed_emulation-style arithmetic on fp32-packed weights and rounds to fp16 only at the sites a scenario
names, then reports max|d|/max|ref| and rel-L2 against the all-fp32 run.  Sites: "<block>.h", "<block>.wp" (conv_0/1/s
weights), "<block>.actv", "<block>.gbw", "<block>.weff", and the globals "codes", "fcmu", "mu", "fc", "xlast", "wimg".
"""
import argparse
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from ctrlhair_b200 import synth  # noqa: E402
from ctrlhair_b200.packer import BLOCKS, ace_list, pack_generator  # noqa: E402
from _packed_emulation import _unpack, _untile  # noqa: E402


def emulate_sites(packed, labels, codes, noise_planes, rounded, ngf=64, label_nc=19):
    def r(site, t):
        return t.half().float() if rounded(site) else t

    B, S = labels.shape[0], labels.shape[1]
    sw = S // 32
    onehot_full = F.one_hot(labels.long(), 32).permute(0, 3, 1, 2).float()

    def onehot_at(res):
        step = S // res
        return onehot_full[:, :, ::step, ::step]

    codes16 = r("codes", codes.float())
    fw, fb = r("fcmu", packed["fcmu.w"].float()), packed["fcmu.b"].float()
    L = codes.shape[2]
    mu_all = r("mu", F.relu(torch.einsum("jnk,bjk->bjn", fw, codes16) + fb[None]))
    noise = iter(noise_planes if noise_planes is not None else [None] * 18)
    x = F.conv2d(onehot_at(sw), _unpack(r("fc", packed["fc.w"].float()), 32), packed["fc.b"], padding=1)
    style_idx = 0
    mults = (1, 2, 2, 4, 8, 16, 32)
    prev_r = sw
    for (name, fi, fo, styled), mul in zip(BLOCKS, mults):
        fin, fout = fi * ngf, fo * ngf
        res = sw * mul
        if res != prev_r:
            x = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
        prev_r = res
        aces = ace_list(fin, fout)
        oh = onehot_at(res)
        actv_all = r(name + ".actv", F.relu(F.conv2d(oh, _unpack(r(name + ".shw", packed[name + ".sh.w"].float()), 32),
                                                     packed[name + ".sh.b"], padding=1)))
        hs = {}

        def modulate(ai, xin, act, tag):
            nonlocal style_idx
            a, C = aces[ai]
            p = "%s.%s" % (name, a)
            bn = min(256, 2 * C)
            actv = actv_all[:, 128 * ai:128 * (ai + 1)]
            gb = F.conv2d(actv, _unpack(r(name + ".gbw", packed[p + ".gb.w"].float()), 128), packed[p + ".gb.b"],
                          padding=1)
            if styled:
                mu = mu_all[:, :, style_idx * L:(style_idx + 1) * L]
                style_idx += 1
                weff = r(name + ".weff", torch.einsum("nk,bjk->bnj", r("stylew", packed[p + ".style.w"].float()), mu))
                weff = weff.reshape(B, 2 * C, 3, 3, label_nc).permute(0, 1, 4, 2, 3)
                gb = gb + torch.cat([F.conv2d(oh[b:b + 1, :label_nc], weff[b], padding=1) for b in range(B)])
            g, be = _untile(gb, bn)
            chan = packed[p + ".chan"]
            xn = xin * chan[0][None, :, None, None] + chan[1][None, :, None, None]
            nz = next(noise)
            if nz is not None:
                xn = xn + nz[..., 0].transpose(1, 2)[:, None] * chan[2][None, :, None, None]
            h = xn * (1 + g) + be
            if act:
                h = F.leaky_relu(h, 0.2)
            return r(name + (".h%s" % tag), h)

        def wp(key, C, taps=9):
            return _unpack(r(name + ".w" + key[6], packed[name + key].float()), C, taps)

        ai = 0
        if fin != fout:
            hs["s"] = modulate(0, x, False, "s")
            ai = 1
        h0 = modulate(ai, x, True, "0")
        dx0 = F.conv2d(h0, wp(".conv_0.w", fin), packed[name + ".conv_0.b"], padding=1)
        h1 = modulate(ai + 1, dx0, True, "1")
        out = F.conv2d(h1, wp(".conv_1.w", min(fin, fout)), packed[name + ".conv_1.b"], padding=1)
        if fin != fout:
            out = out + F.conv2d(hs["s"], wp(".conv_s.w", fin, 1))
        else:
            out = out + x
        x = out
    x = r("xlast", F.leaky_relu(x, 0.2))
    img = F.conv2d(x, _unpack(r("wimg", packed["conv_img.w"].float()), ngf)[:3], packed["conv_img.b"][:3], padding=1)
    return torch.tanh(img)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--dist", default="blocky")
    ap.add_argument("--crop", type=int, default=256)
    ap.add_argument("--exact", action="append", default=[],
                    help="scenario: comma list of sites kept exact, e.g. X,hs@*,ws@*,h1@up_3 (X = xlast+wimg)")
    a = ap.parse_args()
    torch.manual_seed(0)
    sd = synth.make_state_dict()
    packed = pack_generator(sd, weight_dtype=torch.float32)
    del sd
    labels = synth.make_labels(a.B, a.crop, a.dist)
    codes = synth.make_codes(a.B)
    noise = synth.make_noise_planes(a.B, a.crop) if hasattr(synth, "make_noise_planes") else None
    with torch.no_grad():
        t0 = time.time()
        ref = emulate_sites(packed, labels, codes, noise, lambda s: False)
        print("fp32 run %.1fs, max|ref| %.3f" % (time.time() - t0, float(ref.abs().max())), flush=True)
        names = [b[0] for b in BLOCKS]
        late2 = set(names[-2:])
        late1 = set(names[-1:])
        late3 = set(names[-3:])

        def blk(s):
            return s.split(".")[0] if "." in s else None

        X = ("xlast", "wimg")
        H = (".hs", ".h0", ".h1")
        Wt = (".ws", ".w0", ".w1")

        def ex(sites, blocks):
            return lambda s: s not in X and not (s.endswith(sites) and blk(s) in blocks)

        allb = set(names)
        scen = {
            "all fp16 (current)": lambda s: True,
            "exact X": lambda s: s not in X,
            "exact X, up_3 hs+ws": ex((".hs", ".ws"), late1),
            "exact X, up_3 h0+w0": ex((".h0", ".w0"), late1),
            "exact X, up_3 h1+w1": ex((".h1", ".w1"), late1),
            "exact X, up_3 hs+ws+h1+w1": ex((".hs", ".ws", ".h1", ".w1"), late1),
            "exact X, up_3 all": ex(H + Wt, late1),
            "exact X, all hs+ws": ex((".hs", ".ws"), allb),
            "exact X, all hs+ws+h1+w1": ex((".hs", ".ws", ".h1", ".w1"), allb),
            "exact X, up123 hs+ws+h1+w1": ex((".hs", ".ws", ".h1", ".w1"), late3),
            "exact X, up123 hs+ws, up_3 h1+w1": lambda s: ex((".hs", ".ws"), late3)(s) and ex((".h1", ".w1"), late1)(s),
            "exact X, up123 all": ex(H + Wt, late3),
            "exact X, all all": ex(H + Wt, allb),
        }
        if a.exact:
            scen = {}
            for spec in a.exact:
                items = spec.split(",")

                def f(s, items=items):
                    for it in items:
                        if it == "X":
                            if s in X:
                                return False
                        elif "@" in it:
                            st, b = it.split("@")
                            if s.endswith("." + st) and (b == "*" or blk(s) == b):
                                return False
                        elif s == it:
                            return False
                    return True
                scen[spec] = f
        for k, f in scen.items():
            t0 = time.time()
            out = emulate_sites(packed, labels, codes, noise, f)
            d = out - ref
            print("%-44s max-norm %.3e  rel-L2 %.3e  (%.0fs)" % (k, float(d.abs().max() / ref.abs().max()),
                                                             float(d.norm() / ref.norm()), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
