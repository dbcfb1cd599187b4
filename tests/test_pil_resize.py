"""Pillow's BILINEAR resize (my_parsing_util.py:35 resizes every face to 512x512 before parsing), restated bit exactly:
the coefficient tables (ctrlhair_b200.bisenet.pil_bilinear_tables) against Pillow itself on the CPU, and the CUDA passes
(chb_pil_resize_bilinear) against Pillow on the GPU box.  Integer work: the bar is equality."""
import numpy as np
import pytest
import torch
from PIL import Image

from ctrlhair_b200.bisenet import pil_bilinear_tables

SIZES = [((256, 256), (512, 512)), ((300, 280), (512, 512)), ((700, 1024), (512, 512)), ((64, 96), (80, 50)),
         ((512, 512), (512, 512))]


def _numpy_resize(img, out_h, out_w):
    def one(a, O, axis):
        a = np.moveaxis(a, axis, 0).astype(np.int64)
        b, c, _ = pil_bilinear_tables(a.shape[0], O)
        out = np.zeros((O,) + a.shape[1:], np.int64)
        for o in range(O):
            acc = np.full(a.shape[1:], 1 << 21, np.int64)
            for k in range(b[o, 1]):
                acc = acc + a[b[o, 0] + k] * int(c[o, k])
            out[o] = np.clip(acc >> 22, 0, 255)
        return np.moveaxis(out, 0, axis).astype(np.uint8)
    return one(one(img, out_w, 1), out_h, 0)


@pytest.mark.parametrize("hw,out", SIZES)
def test_tables_reproduce_pillow(hw, out):
    rng = np.random.default_rng(hw[0] * 7 + out[0])
    img = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
    want = np.asarray(Image.fromarray(img).resize((out[1], out[0]), Image.BILINEAR))
    assert np.array_equal(_numpy_resize(img, out[0], out[1]), want)


@pytest.mark.gpu
@pytest.mark.parametrize("hw,out", SIZES)
def test_cuda_resize_equals_pillow(hw, out):
    from ctrlhair_b200.bisenet import resize_bilinear_u8
    rng = np.random.default_rng(hw[1] * 3 + out[1])
    imgs = rng.integers(0, 256, (3,) + hw + (3,), dtype=np.uint8)
    imgs[1] = 255
    imgs[2, ::2] = 0
    got = resize_bilinear_u8(torch.from_numpy(imgs).cuda(), out[0], out[1]).cpu().numpy()
    for i in range(3):
        want = np.asarray(Image.fromarray(imgs[i]).resize((out[1], out[0]), Image.BILINEAR))
        assert np.array_equal(got[i], want), i
