"""TEST INFRASTRUCTURE — CPU oracle of the shape branch generator (mask encoders / decoders), restated from:

  positional embedding   shape_branch/model.py:18-30
  MaskEncoder            shape_branch/model.py:69-113   (7 x [ZeroPad(1), conv4x4 s2, LayerNorm, LeakyReLU 0.2] -> fc)
  MaskDecoder            shape_branch/model.py:116-143  (fc -> 7 x [nearest up2, ZeroPad(1), conv3x3, LayerNorm, LeakyReLU] -> conv3x3)
  Generator.forward_*    shape_branch/model.py:164-199  (hair logit inserted at channel 13, softmax over 19)
  Conv2dBlock / LinearBlock / LayerNorm   my_torchlib/module.py:129-137, 56-64, 189-205
      LayerNorm: per-sample mean and UNBIASED std over (C,H,W), x = (x - mean) / (std + 1e-5) * gamma_c + beta_c
Pinned by tests/golden/shape_*.npz (outputs of the unmodified reference, oracle/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F

HAIR_IDX = 13  # global_value_utils.py:49-52


def pos_embedding(img_size=256, order=10):
    c = torch.arange(img_size, dtype=torch.float64) / img_size       # np.linspace(0, 1, n, endpoint=False)
    gx, gy = torch.meshgrid(c, c, indexing="xy")                       # np.meshgrid default
    bi = torch.stack([gx, gy], 0)[None]                                # [1,2,H,W]
    nums = (2.0 ** torch.arange(order, dtype=torch.float64) * math.pi)[:, None, None, None]
    gamma = torch.cat([torch.sin(nums * bi), torch.cos(nums * bi)], 0)  # [2*order, 2, H, W]
    return gamma.reshape(-1, img_size, img_size).float()                 # [4*order, H, W]


def layer_norm(x, gamma, beta, eps=1e-5):
    flat = x.reshape(x.shape[0], -1)
    mean = flat.mean(1).view(-1, 1, 1, 1)
    std = flat.std(1).view(-1, 1, 1, 1)       # unbiased
    x = (x - mean) / (std + eps)
    return x * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def mask_encoder(sd, p, mask, n_layers=7, vae=False):
    """mask fp32 [B, 1|18, 256, 256] -> mean code (and |std| for the VAE hair encoder)."""
    B = mask.shape[0]
    x = torch.cat([mask, pos_embedding(mask.shape[2])[None].expand(B, -1, -1, -1)], 1)
    for i in range(n_layers):
        q = "%s.layers.%d." % (p, i)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[q + "conv.weight"], sd[q + "conv.bias"], stride=2)
        x = F.leaky_relu(layer_norm(x, sd[q + "norm.gamma"], sd[q + "norm.beta"]), 0.2)
    feat = x.flatten(1)
    mean = F.linear(feat, sd[p + ".out_layer.fc.weight"], sd[p + ".out_layer.fc.bias"])
    if vae:
        std = F.linear(feat, sd[p + ".std_out_layer.fc.weight"], sd[p + ".std_out_layer.fc.bias"]).abs()
        return mean, std
    return mean, None


def mask_decoder(sd, p, code, n_layers=7):
    x = F.linear(code, sd[p + ".in_layer.fc.weight"], sd[p + ".in_layer.fc.bias"])
    x = x.reshape(-1, 2048, 2, 2)
    for i in range(n_layers):
        q = "%s.layers.%d." % (p, 2 * i + 1)
        x = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[q + "conv.weight"], sd[q + "conv.bias"])
        x = F.leaky_relu(layer_norm(x, sd[q + "norm.gamma"], sd[q + "norm.beta"]), 0.2)
    return F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[p + ".out_layer.conv.weight"], sd[p + ".out_layer.conv.bias"])


def forward_hair_encoder(sd, hair, testing=True):
    mean, std = mask_encoder(sd, "hair_encoder", hair, vae=True)
    return mean if testing else (mean, std)


def forward_face_encoder(sd, face):
    return mask_encoder(sd, "face_encoder", face)[0]


def forward_decoder(hair_logit, face_logit):
    logit = torch.cat([face_logit[:, :HAIR_IDX], hair_logit, face_logit[:, HAIR_IDX:]], 1)
    return torch.softmax(logit, 1)


def forward_decode_by_code(sd, hair_code, face_code):
    hair_logit = mask_decoder(sd, "hair_decoder", torch.cat([face_code, hair_code], 1))
    face_logit = mask_decoder(sd, "face_decoder", face_code)
    return forward_decoder(hair_logit, face_logit)
