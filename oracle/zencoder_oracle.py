"""TEST INFRASTRUCTURE — CPU oracle of the SEAN style encoder (Zencoder), restated from the reference:

  network      sean_codes/models/networks/architecture.py:154-175  (ReflPad+conv3x3, 2x conv s2, convT s2, ReflPad+conv3x3+tanh,
               each but the last followed by InstanceNorm2d(affine=False, eps 1e-5) + LeakyReLU(0.2))
  region pool  architecture.py:177-207  (seg nearest-resized to the code map; per (image, class) with area > 0 the mean of
               the 512-channel code over the region; absent classes stay zero)
Pinned by tests/golden/zencoder_*.npz (outputs of the unmodified reference module, oracle/make_golden.py).
"""
import torch
import torch.nn.functional as F

from .sean_oracle import nearest, one_hot


def _in_lrelu(x):
    return F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)


def code_map(sd, img):
    p = "Zencoder.model."
    x = F.conv2d(F.pad(img, (1, 1, 1, 1), mode="reflect"), sd[p + "1.weight"], sd[p + "1.bias"])
    x = _in_lrelu(x)
    x = _in_lrelu(F.conv2d(x, sd[p + "4.weight"], sd[p + "4.bias"], stride=2, padding=1))
    x = _in_lrelu(F.conv2d(x, sd[p + "7.weight"], sd[p + "7.bias"], stride=2, padding=1))
    x = _in_lrelu(F.conv_transpose2d(x, sd[p + "10.weight"], sd[p + "10.bias"], stride=2, padding=1, output_padding=1))
    x = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), sd[p + "14.weight"], sd[p + "14.bias"])
    return torch.tanh(x)


def region_pool(codes, seg):
    B, C = codes.shape[0], codes.shape[1]
    seg = nearest(seg, codes.shape[2])
    out = torch.zeros((B, seg.shape[1], C), dtype=codes.dtype)
    for i in range(B):
        for j in range(seg.shape[1]):
            mask = seg[i, j].bool()
            area = int(mask.sum())
            if area > 0:
                out[i, j] = codes[i][:, mask].mean(1)
    return out


def zencoder_forward(sd, img, labels):
    """img fp32 [B,3,S,S] in [-1,1]; labels integer [B,S,S] -> style codes [B,19,512]."""
    return region_pool(code_map(sd, img), one_hot(labels))
