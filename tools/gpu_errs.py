"""Measured errors of the secondary networks against their oracles / goldens (development aid): python tools/gpu_errs.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctrlhair_b200 import synth  # noqa: E402
from oracle import shape_oracle as sho  # noqa: E402
from oracle import zencoder_oracle as zo  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def e(got, want):
    d = got - want
    return "rel-L2 %.2e  max-norm %.2e" % (float(d.norm() / want.norm()), float(d.abs().max() / want.abs().max()))


def main():
    from ctrlhair_b200.shape import ShapeGeneratorB200
    from ctrlhair_b200.zencoder import ZencoderB200
    sd = synth.make_state_dict()
    for name, S, B in (("zencoder_c64_b2", 64, 2), ("zencoder_c256_b1", 256, 1)):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        img, labels, gold = synth.make_image(B, S), torch.from_numpy(g["labels"]), torch.from_numpy(g["out"])
        enc = ZencoderB200(crop=S, max_batch=B).load_state_dict(sd)
        out = enc(img.cuda(), labels.cuda()).cpu()
        print(name, "vs golden:", e(out, gold), " vs oracle:", e(out, zo.zencoder_forward(sd, img, labels)))
    ssd = synth.make_shape_state_dict()
    g = np.load(os.path.join(GOLD, "shape_b2.npz"))
    net = ShapeGeneratorB200(max_batch=2).load_state_dict(ssd)
    hair, face = synth.make_shape_inputs(2)
    hc = net.forward_hair_encoder(hair.cuda(), testing=True).cpu()
    fc = net.forward_face_encoder(face.cuda()).cpu()
    rh, rf = torch.from_numpy(g["hair_code"]), torch.from_numpy(g["face_code"])
    print("shape hair code:", e(hc, rh), " face code:", e(fc, rf))
    m = net.forward_decode_by_code(rh.cuda(), rf.cuda()).cpu()
    ref = sho.forward_decode_by_code(ssd, rh, rf)
    print("shape decoder probs: max abs %.2e  argmax agreement %.5f" %
          (float((m - ref).abs().max()), float((m.argmax(1) == ref.argmax(1)).float().mean())))
    hl = net.forward_hair_decoder(rh.cuda(), rf.cuda()).cpu()
    fl = net.forward_face_decoder(rf.cuda()).cpu()
    print("shape hair logits:", e(hl, sho.mask_decoder(ssd, "hair_decoder", torch.cat([rf, rh], 1))),
          " face logits:", e(fl, sho.mask_decoder(ssd, "face_decoder", rf)))


if __name__ == "__main__":
    main()
