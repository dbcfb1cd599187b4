"""BiSeNet face parsing on the GPU (csrc/bisenet.cu) against the golden outputs of the UNMODIFIED
reference network (tests/golden/bisenet_b1.npz, oracle/make_golden_bisenet.py) and against the oracle.

Tolerances: the label maps are integer results of an argmax over fp16-operand logits, so they are compared by
agreement: >= 99.9 % of the pixels, and every disagreeing pixel must be a near-tie of the fp32 logits (top-2 margin
below the logit error bound).  1/8-resolution logits: max|d| / max|ref| <= 3e-3 (28 fp16-operand convs deep)."""
import os

import numpy as np
import pytest
import torch

from ctrlhair_b200 import _lib, synth
from oracle import bisenet_oracle as bno

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bisenet_b1.npz")


@pytest.fixture(scope="module")
def nets():
    from ctrlhair_b200.bisenet import BiSeNetB200
    sd = synth.make_bisenet_state_dict()
    raw = BiSeNetB200(max_batch=2, swap_labels=False).load_state_dict(sd)
    swapped = BiSeNetB200(max_batch=2, swap_labels=True).load_state_dict(sd)
    return sd, raw, swapped


def test_parsing_vs_reference_golden(nets):
    sd, raw, swapped = nets
    g = np.load(GOLD)
    img = torch.from_numpy(g["img"]).cuda()
    parsing, logits = raw(img, return_logits=True)
    with torch.no_grad():
        ref_low = bno.bisenet_logits_lowres(sd, bno.normalise_image(g["img"]))           # [1,19,64,64]
        ref_full = torch.nn.functional.interpolate(ref_low, (512, 512), mode="bilinear", align_corners=True)
    got_low = logits.cpu().permute(0, 3, 1, 2)
    err = float((got_low - ref_low).abs().max())
    scale = float(ref_low.abs().max())
    assert torch.isfinite(got_low).all() and err <= 3e-3 * scale, (err, scale)
    # the golden file pins the oracle's upsampled logits to the reference's own
    assert float((ref_full[:, :, ::16, ::16] - torch.from_numpy(g["logits_sub"])).abs().max()) < 2e-5 * scale
    got = parsing[0].cpu().numpy()
    want = g["parsing"]
    agree = float((got == want).mean())
    top2 = ref_full[0].topk(2, dim=0).values
    margin = (top2[0] - top2[1]).numpy()
    bad = got != want
    # every disagreement sits on a near-tie of the reference logits: margin below twice the measured logit error
    assert agree >= 0.999, agree
    assert not bad.any() or float(margin[bad].max()) <= 2.0 * err + 1e-6, (float(margin[bad].max()), err)
    print("bisenet: label agreement %.5f, logit max err %.2e of %.2f, median top-2 margin %.3f, %d classes" %
          (agree, err, scale, float(np.median(margin)), len(np.unique(want))))


def test_get_mask_swaps_labels_and_resizes(nets):
    sd, raw, swapped = nets
    g = np.load(GOLD)
    mask = swapped.get_mask(g["img"][0], img_size=256)          # already 512x512: the PIL resize is the identity
    assert mask.shape == (256, 256) and mask.dtype == np.uint8
    assert float((mask == g["mask256"]).mean()) >= 0.999
    # swap == LUT applied to the network-order map, resize == every second pixel (cv2 INTER_NEAREST 512 -> 256)
    full = raw(torch.from_numpy(g["img"]).cuda())[0].cpu().numpy()
    assert np.array_equal(mask, bno.swap_parsing_label_to_celeba_mask(full)[::2, ::2].astype(np.uint8))


def test_batch_and_host_entry_points(nets):
    sd, raw, swapped = nets
    g = np.load(GOLD)
    rng = np.random.default_rng(3)
    img2 = np.stack([g["img"][0], rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)])
    a = swapped(torch.from_numpy(img2).cuda(), out_size=256).cpu()
    b = swapped.forward_host(img2, out_size=256)
    assert torch.equal(a, b)
    solo = swapped(torch.from_numpy(img2[1:2]).cuda(), out_size=256).cpu()
    assert torch.equal(solo[0], a[1])                            # an image's result does not depend on its batch-mates
    with pytest.raises(_lib.ChbError):
        swapped(torch.zeros((3, 512, 512, 3), dtype=torch.uint8).cuda())      # exceeds max_batch
    with pytest.raises(_lib.ChbError):
        swapped(torch.zeros((1, 256, 256, 3), dtype=torch.uint8).cuda())      # not the network size


def test_get_mask_resizes_like_pil_on_the_device(nets):
    """get_mask on a 256x256 face (what Backend.parse_img hands over, ui/backend.py:69-74) == parsing the image Pillow
    resizes to 512x512 on the host (my_parsing_util.py:33-35): the device resize is bit exact, so the masks are equal."""
    sd, raw, swapped = nets
    g = np.load(GOLD)
    small = np.ascontiguousarray(g["img"][0, ::2, ::2])                       # 256x256x3
    big = swapped.resize_to_network(small, 512)                              # Pillow on the host
    want = swapped(torch.from_numpy(big[None].copy()).cuda(), out_size=256)[0].cpu().numpy()
    got = swapped.get_mask(small, img_size=256)
    assert got.shape == (256, 256) and np.array_equal(got, want)
    batch = swapped.get_mask(np.stack([small, small[::-1].copy()]), img_size=256)
    assert np.array_equal(batch[0], want)
