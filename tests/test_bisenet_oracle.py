"""Groundwork for SURVEY §8f row 3 (face parsing on the GPU; no CUDA path yet): the BiSeNet oracle is pinned to the
unmodified reference through tests/golden/bisenet_b1.npz (oracle/make_golden_bisenet.py), and the conv + BatchNorm
folding a weight packer will apply is checked against it on the CPU."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from ctrlhair_b200 import synth
from oracle import bisenet_oracle as bno

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bisenet_b1.npz")


def test_bisenet_oracle_matches_reference_golden():
    g = np.load(GOLD)
    sd = synth.make_bisenet_state_dict()
    with torch.no_grad():
        logits = bno.bisenet_forward(sd, bno.normalise_image(g["img"]))
    want = torch.from_numpy(g["logits_sub"])
    assert float((logits[:, :, ::16, ::16] - want).abs().max()) < 2e-5 * float(want.abs().max())
    parsing = bno.parsing_labels(logits)[0]
    assert float((parsing == g["parsing"]).mean()) > 0.9999       # argmax ties at fp32 rounding distance
    mask = bno.get_mask(sd, g["img"])[0]
    assert mask.shape == (256, 256) and float((mask == g["mask256"]).mean()) > 0.9999


def test_label_swap_is_the_reference_loop():
    parsing = np.arange(19).repeat(3).reshape(3, 19)
    want = np.zeros_like(parsing)
    for label_idx, label_name in enumerate(bno.PARSING_LABEL_LIST):          # my_parsing_util.py:49-54
        want[bno.BISENET_LABELS.index(label_name) == parsing] = label_idx
    assert np.array_equal(bno.swap_parsing_label_to_celeba_mask(parsing), want)
    assert sorted(bno.BISENET_LABELS) == sorted(bno.PARSING_LABEL_LIST)


def test_conv_bn_folding_reproduces_the_oracle():
    """Eval BatchNorm folds into the preceding bias-free conv: W' = W * g / sqrt(var + eps), b' = beta - mean * g /
    sqrt(var + eps).  Every conv of the net but the four attention / output 1x1s is followed by one."""
    sd = synth.make_bisenet_state_dict()
    gen = torch.Generator().manual_seed(3)
    for conv, bn, stride, pad in [("cp.resnet.conv1", "cp.resnet.bn1", 2, 3),
                                  ("cp.resnet.layer2.0.conv1", "cp.resnet.layer2.0.bn1", 2, 1),
                                  ("cp.resnet.layer2.0.downsample.0", "cp.resnet.layer2.0.downsample.1", 2, 0),
                                  ("cp.arm32.conv.conv", "cp.arm32.conv.bn", 1, 1),
                                  ("ffm.convblk.conv", "ffm.convblk.bn", 1, 0)]:
        w = sd[conv + ".weight"]
        x = torch.randn((2, w.shape[1], 12, 12), generator=gen)
        want = F.batch_norm(F.conv2d(x, w, None, stride, pad), sd[bn + ".running_mean"], sd[bn + ".running_var"],
                            sd[bn + ".weight"], sd[bn + ".bias"], False, 0.1, bno.BN_EPS)
        s = sd[bn + ".weight"] * torch.rsqrt(sd[bn + ".running_var"] + bno.BN_EPS)
        got = F.conv2d(x, w * s[:, None, None, None], sd[bn + ".bias"] - sd[bn + ".running_mean"] * s, stride, pad)
        assert float((got - want).abs().max()) < 1e-5 * float(want.abs().max()), conv


def test_packed_schedule_is_exact_in_fp32():
    """The blob the CUDA path consumes (BatchNorm folded, identity shortcuts as identity 1x1 K-segments, the FFM concat as
    two K-segments, stride-2 convs at full resolution + phase pick) reproduces the oracle in fp32; with fp16 rounding at
    the kernels' storage points the labels still agree on > 99.9 % of the pixels."""
    import _bisenet_emulation as be
    from ctrlhair_b200.bisenet import label_lut, pack_bisenet
    g = np.load(GOLD)
    sd = synth.make_bisenet_state_dict()
    packed = {k: (v.float() if v.dtype == torch.float16 else v) for k, v in pack_bisenet(sd).items()}
    img = g["img"][:, :256, :256]            # a 256x256 crop keeps the CPU test short; the network is fully convolutional
    with torch.no_grad():
        ref = bno.bisenet_logits_lowres(sd, bno.normalise_image(img))
        got = be.emulate(packed, img)
        got16 = be.emulate(packed, img, round16=True)
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) < 2e-3 * scale      # fp16 WEIGHTS only (activations fp32)
    up = lambda t: F.interpolate(t, (256, 256), mode="bilinear", align_corners=True).argmax(1)
    assert float((up(got16) == up(ref)).float().mean()) > 0.999
    lut = label_lut(True)
    assert np.array_equal(lut, bno.swap_parsing_label_to_celeba_mask(np.arange(19)))
