"""One-hot pyramid (integer work: bit exact against the oracle) and the device noise generator."""
import pytest
import torch

from ctrlhair_b200 import ops
from oracle import sean_oracle as so
from ctrlhair_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["iid", "blocky"])
@pytest.mark.parametrize("S,B", [(256, 3), (64, 2), (32, 1)])
def test_onehot_pyramid_bit_exact(kind, S, B):
    labels = synth.make_labels(B, S, kind)
    res = [S >> s for s in range(5, -1, -1)]
    outs = ops.onehot_pyramid(labels.cuda(), res)
    oh = so.one_hot(labels)  # [B,19,S,S] fp32, the reference scatter_
    for r, o in zip(res, outs):
        want = so.nearest(oh, r).permute(0, 2, 3, 1)  # NHWC
        got = o.cpu().float()
        assert got.shape == (B, r, r, 32)
        assert torch.equal(got[..., :19], want)
        assert torch.count_nonzero(got[..., 19:]) == 0


def test_onehot_out_of_range_label_is_all_zero():
    labels = torch.zeros((1, 32, 32), dtype=torch.uint8)
    labels[0, 0, 0] = 255  # the reference would raise in scatter_; here the pixel simply has no class
    labels[0, 1, 1] = 18
    (o,) = ops.onehot_pyramid(labels.cuda(), [32])
    o = o.cpu()
    assert torch.count_nonzero(o[0, 0, 0]) == 0
    assert o[0, 1, 1, 18] == 1 and torch.count_nonzero(o[0, 1, 1]) == 1


def test_noise_fill_statistics_and_determinism():
    n = 1 << 22
    a = ops.noise_fill(n, seed=1237)
    b = ops.noise_fill(n, seed=1237)
    c = ops.noise_fill(n, seed=1238)
    assert torch.equal(a, b) and not torch.equal(a, c)
    assert abs(float(a.mean())) < 3e-3 and abs(float(a.var()) - 1.0) < 5e-3
    assert torch.isfinite(a).all() and float(a.abs().max()) < 7.0
    # kurtosis of a normal is 3
    assert abs(float((a ** 4).mean()) - 3.0) < 0.05
    # counter offset continues the same stream (4 normals per counter)
    tail = ops.noise_fill(n - 1024, seed=1237, offset=256)
    assert torch.equal(tail, a[1024:])
