"""TEST INFRASTRUCTURE ONLY — CPU restatement (functional torch, fp32) of the reference's face-parsing network, the
checker of SURVEY §8f row 3 (BiSeNet parsing on the GPU: csrc/bisenet.cu, ctrlhair_b200/bisenet.py).  Nothing in
ctrlhair_b200/ uses it.  Only tests/ may import it.

  bisenet_forward      external_code/face_parsing/model.py:257-274 (first head only: parsing_img uses out[0]),
                       ContextPath :116-146, AttentionRefinementModule :81-103, FeatureFusionModule :196-231,
                       BiSeNetOutput :37-56, ConvBNReLU :12-35; external_code/face_parsing/resnet.py:21-93
  normalise_image      my_parsing_util.py:25-28,35-36 (ToTensor + Normalize) for an image that is already 512x512
                       (the PIL bilinear resize before it is third-party Pillow code and stays on the host)
  parsing_labels       my_parsing_util.py:45-46 (argmax over the 19 logits)
  swap_parsing_label_to_celeba_mask   my_parsing_util.py:49-54
  get_mask             hair_editor.py:331-335 (label swap + cv2 INTER_NEAREST resize to img_size; for the 512 -> 256
                       case cv2 takes source index floor(dst * 2) = every second pixel)

Parity pin: oracle/make_golden_bisenet.py builds the unmodified reference BiSeNet (torch.utils.model_zoo.load_url
stubbed: the ImageNet ResNet-18 download is overwritten by load_state_dict anyway), loads the same synthetic checkpoint
and stores its outputs in tests/golden/bisenet_b1.npz; tests/test_bisenet_oracle.py checks this file against them.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5
# my_parsing_util.py:18-22 (network's label order) and global_value_utils.py:49-51 (CelebAMask-HQ order used downstream)
BISENET_LABELS = ["background", "skin_other", "l_brow", "r_brow", "l_eye", "r_eye", "eye_g", "l_ear", "r_ear", "ear_r",
                  "nose", "mouth", "u_lip", "l_lip", "neck", "neck_l", "cloth", "hair", "hat"]
PARSING_LABEL_LIST = ["background", "skin_other", "nose", "eye_g", "l_eye", "r_eye", "l_brow", "r_brow", "l_ear", "r_ear",
                      "mouth", "u_lip", "l_lip", "hair", "hat", "ear_r", "neck_l", "neck", "cloth"]


def _bn(sd, name, x):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], False, 0.1, BN_EPS)


def _cbr(sd, name, x, stride=1, padding=1):
    """ConvBNReLU (model.py:12-29)."""
    return F.relu(_bn(sd, name + ".bn", F.conv2d(x, sd[name + ".conv.weight"], None, stride, padding)))


def _basic_block(sd, p, x, stride):
    """resnet.py:21-52."""
    r = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, stride, 1)))
    r = _bn(sd, p + ".bn2", F.conv2d(r, sd[p + ".conv2.weight"], None, 1, 1))
    s = x
    if (p + ".downsample.0.weight") in sd:
        s = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], None, stride, 0))
    return F.relu(s + r)


def resnet18(sd, x, p="cp.resnet"):
    """resnet.py:72-81 -> (feat8, feat16, feat32)."""
    x = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], None, 2, 3)))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for li in range(1, 5):
        for bi in range(2):
            x = _basic_block(sd, "%s.layer%d.%d" % (p, li, bi), x, 2 if (li > 1 and bi == 0) else 1)
        feats.append(x)
    return feats[1], feats[2], feats[3]


def _arm(sd, name, x):
    """AttentionRefinementModule (model.py:91-98)."""
    feat = _cbr(sd, name + ".conv", x)
    atten = F.avg_pool2d(feat, feat.shape[2:])
    atten = torch.sigmoid(_bn(sd, name + ".bn_atten", F.conv2d(atten, sd[name + ".conv_atten.weight"])))
    return feat * atten


def context_path(sd, x):
    """model.py:127-146 -> (feat8, feat16_up, feat32_up)."""
    feat8, feat16, feat32 = resnet18(sd, x)
    avg = F.avg_pool2d(feat32, feat32.shape[2:])
    avg = _cbr(sd, "cp.conv_avg", avg, 1, 0)
    avg_up = F.interpolate(avg, feat32.shape[2:], mode="nearest")
    feat32_sum = _arm(sd, "cp.arm32", feat32) + avg_up
    feat32_up = _cbr(sd, "cp.conv_head32", F.interpolate(feat32_sum, feat16.shape[2:], mode="nearest"))
    feat16_sum = _arm(sd, "cp.arm16", feat16) + feat32_up
    feat16_up = _cbr(sd, "cp.conv_head16", F.interpolate(feat16_sum, feat8.shape[2:], mode="nearest"))
    return feat8, feat16_up, feat32_up


def feature_fusion(sd, fsp, fcp):
    """FeatureFusionModule (model.py:218-228)."""
    feat = _cbr(sd, "ffm.convblk", torch.cat([fsp, fcp], dim=1), 1, 0)
    atten = F.avg_pool2d(feat, feat.shape[2:])
    atten = F.relu(F.conv2d(atten, sd["ffm.conv1.weight"]))
    atten = torch.sigmoid(F.conv2d(atten, sd["ffm.conv2.weight"]))
    return feat * atten + feat


def bisenet_logits_lowres(sd, x):
    """x float [B,3,H,W] (normalised) -> the main head's logits [B,19,H/8,W/8] BEFORE the bilinear upsample."""
    feat_res8, feat_cp8, _ = context_path(sd, x)
    feat_fuse = feature_fusion(sd, feat_res8, feat_cp8)
    return F.conv2d(_cbr(sd, "conv_out.conv", feat_fuse), sd["conv_out.conv_out.weight"])


def bisenet_forward(sd, x):
    """x float [B,3,H,W] (normalised) -> logits [B,19,H,W] of the main head (model.py:257-270, out[0])."""
    H, W = x.shape[2:]
    return F.interpolate(bisenet_logits_lowres(sd, x), (H, W), mode="bilinear", align_corners=True)


def normalise_image(img_u8):
    """uint8 [B,H,W,3] RGB -> float [B,3,H,W]: ToTensor (/255) + Normalize(ImageNet mean / std)."""
    x = torch.as_tensor(np.asarray(img_u8)).permute(0, 3, 1, 2).float() / 255.0
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (x - mean) / std


def parsing_labels(logits):
    """my_parsing_util.py:45-46: argmax over classes (first maximum)."""
    return logits.argmax(1).numpy()


def swap_parsing_label_to_celeba_mask(parsing):
    """my_parsing_util.py:49-54 as a lookup table: network label i -> index of its name in PARSING_LABEL_LIST."""
    lut = np.array([PARSING_LABEL_LIST.index(n) for n in BISENET_LABELS])
    return lut[np.asarray(parsing)]


def get_mask(sd, img512_u8, img_size=256):
    """hair_editor.py:331-335 for 512x512 inputs: parse, swap the labels, nearest-resize to img_size."""
    lab = swap_parsing_label_to_celeba_mask(parsing_labels(bisenet_forward(sd, normalise_image(img512_u8))))
    step = lab.shape[1] // img_size
    return lab[:, ::step, ::step].astype(np.uint8)
