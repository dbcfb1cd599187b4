"""Pins oracle/blend_oracle.py against the unmodified reference and against cv2, and writes tests/golden/blend.npz.

Run in the build container (needs /root/reference and cv2; neither exists on the GPU box):
    python oracle/make_golden_blend.py
* poisson_blending: the reference function (poisson_blending.py:29-87) is imported as it is and run on seeded inputs
  (three small ragged sizes and one 96x96 face-like case); inputs and outputs are stored.
* blend_mask: hair_editor.py:297-306 is a method of a class whose import needs dlib etc.; its five cv2 / numpy lines are
  executed here directly with cv2 (structuring elements and dilations come from cv2 itself).
* RGB<->HSV: cv2.cvtColor over ALL 2^24 RGB triples and all 180x256x256 HSV triples; SHA-256 digests are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CHB_REFERENCE", "/root/reference")

from oracle import blend_oracle as bo  # noqa: E402


def face_like_case(H, W, seed):
    from ctrlhair_b200 import synth
    return synth.make_blend_case(H, W, seed)


def main():
    import cv2
    sys.path.insert(0, REF)
    from poisson_blending import poisson_blending as ref_poisson  # the unmodified reference

    out = {}
    # ---- structuring elements and blend mask vs cv2
    for k in (13, 5):
        se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, ksize=(k, k))
        assert np.array_equal(se, bo.ellipse_kernel(k)), k
    for i, (H, W) in enumerate([(40, 56), (96, 96), (256, 256)]):
        face, gen, fp, tp = face_like_case(H, W, 100 + i)
        res_mask = np.logical_or(tp == bo.HAIR_IDX, fp == bo.HAIR_IDX).astype("uint8")
        k13 = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, ksize=(13, 13))
        k5 = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, ksize=(5, 5))
        d13 = cv2.dilate(res_mask, k13, iterations=1)[..., None]
        d5 = cv2.dilate(res_mask, k5, iterations=1)[..., None]
        bg = (tp[..., None] == 0)
        want = d13 * (1 - bg) + d5 * bg
        got = bo.blend_mask(tp, fp)
        assert np.array_equal(want[..., 0].astype(np.uint8), got), (H, W)
        if (H, W) == (96, 96):
            out["mask_fp"], out["mask_tp"], out["mask_out"] = fp, tp, got
    print("blend_mask == cv2 path on 3 sizes")

    # ---- poisson blending vs the reference function
    g = np.random.default_rng(7)
    cases = []
    for H, W in [(9, 12), (24, 17), (33, 40)]:
        src = g.integers(0, 256, (H, W, 3), dtype=np.uint8)
        tgt = g.integers(0, 256, (H, W, 3), dtype=np.uint8)
        m = (g.random((H, W, 1)) < 0.6).astype(np.uint8)
        m[0, :3] = 0
        m[-1, -2:] = 1                                   # border pixels of both kinds
        cases.append((src, tgt, m))
    face, gen, fp, tp = face_like_case(96, 96, 101)
    rmd = bo.blend_mask(tp, fp)[..., None]
    cases.append((face, gen, (1 - rmd).astype(np.uint8)))
    for i, (src, tgt, m) in enumerate(cases):
        want = ref_poisson(src.copy(), tgt.copy(), m.copy(), with_gamma=True)
        got = bo.poisson_blending(src, tgt, m, with_gamma=True)
        nd = int((want != got).sum())
        print("poisson case %d %s: %d differing bytes of %d" % (i, src.shape, nd, want.size))
        assert nd == 0
        out["p%d_src" % i], out["p%d_tgt" % i], out["p%d_mask" % i], out["p%d_out" % i] = src, tgt, m[..., 0], want
    want = ref_poisson(cases[1][0].copy(), cases[1][1].copy(), cases[1][2].copy(), with_gamma=False)
    assert np.array_equal(want, bo.poisson_blending(*cases[1], with_gamma=False))
    out["p1_out_nogamma"] = want
    out["n_poisson"] = np.array(len(cases))
    out["lut_fwd"], out["lut_known"] = bo.gamma_tables()   # this host's pow, see blend_oracle.gamma_tables

    # ---- colour space vs cv2, exhaustive
    rgb = bo.all_rgb()
    want = cv2.cvtColor(rgb.reshape(-1, 1, 3), cv2.COLOR_RGB2HSV)[:, 0]   # one pixel per row: the scalar path
    got = bo.rgb_to_hsv_u8(rgb)
    assert np.array_equal(want, got), int((want != got).any(-1).sum())
    out["rgb2hsv_sha256"] = np.array(bo.table_digest(want))
    hsv = rgb[rgb[:, 0] < 180]
    want = cv2.cvtColor(hsv.reshape(-1, 1, 3), cv2.COLOR_HSV2RGB)[:, 0]
    got = bo.hsv_to_rgb_u8(hsv)
    nd = int((want != got).any(-1).sum())
    print("HSV2RGB: %d of %d triples differ" % (nd, len(hsv)))
    assert nd == 0
    out["hsv2rgb_sha256"] = np.array(bo.table_digest(want))
    sub = np.random.default_rng(3).integers(0, len(hsv), 4096)
    out["hsv_sample"], out["hsv_sample_rgb"] = hsv[sub], want[sub]
    print("RGB2HSV / HSV2RGB == cv2 on every 8-bit input")

    path = os.path.join(ROOT, "tests", "golden", "blend.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
