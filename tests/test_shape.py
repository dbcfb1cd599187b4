"""Shape branch nets: oracle vs the reference's golden vectors (CPU), CUDA vs oracle/golden (GPU)."""
import os

import numpy as np
import pytest
import torch

from ctrlhair_b200 import synth
from oracle import shape_oracle as sho

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shape_b2.npz")


@pytest.fixture(scope="module")
def shape_sd():
    return synth.make_shape_state_dict()


def test_shape_oracle_matches_reference_golden(shape_sd):
    g = np.load(GOLD)
    hair, face = synth.make_shape_inputs(2)
    hc, fc = sho.forward_hair_encoder(shape_sd, hair), sho.forward_face_encoder(shape_sd, face)
    m = sho.forward_decode_by_code(shape_sd, hc, fc)
    assert float((hc - torch.from_numpy(g["hair_code"])).abs().max()) < 2e-5
    assert float((fc - torch.from_numpy(g["face_code"])).abs().max()) < 2e-5
    assert float((m[:, :, ::8, ::8] - torch.from_numpy(g["mask_sub"])).abs().max()) < 2e-5
    assert float((m.sum(1) - 1).abs().max()) < 1e-5
    assert np.array_equal(m.argmax(1).to(torch.uint8).numpy(), g["mask_argmax"])


def test_positional_embedding_matches_numpy_formula():
    """shape_branch/model.py:18-30 restated; checked against the same numpy expression."""
    n, order = 16, 3
    c = np.linspace(0, 1, n, endpoint=False)
    bi = np.stack(np.meshgrid(c, c), 0)[None]
    nums = (2 ** np.arange(order) * np.pi)[:, None, None, None]
    ref = np.concatenate([np.sin(nums * bi), np.cos(nums * bi)], 0).reshape(-1, n, n).astype(np.float32)
    assert np.allclose(sho.pos_embedding(n, order).numpy(), ref, atol=1e-6)


@pytest.mark.gpu
def test_shape_cuda_matches_oracle_and_golden(shape_sd):
    from ctrlhair_b200 import _lib
    from ctrlhair_b200.shape import ShapeGeneratorB200
    g = np.load(GOLD)
    net = ShapeGeneratorB200(max_batch=2).load_state_dict(shape_sd)
    hair, face = synth.make_shape_inputs(2)
    hc = net.forward_hair_encoder(hair.cuda(), testing=True).cpu()
    fc = net.forward_face_encoder(face.cuda()).cpu()
    rh, rf = torch.from_numpy(g["hair_code"]), torch.from_numpy(g["face_code"])
    # fp16 tensor-core operands through 7 LayerNorm'd conv layers: 1e-3 relative (measured 7.6-7.7e-4)
    assert float((hc - rh).norm() / rh.norm()) < 1e-3, float((hc - rh).norm() / rh.norm())
    assert float((fc - rf).norm() / rf.norm()) < 1e-3, float((fc - rf).norm() / rf.norm())
    # decode from the *reference* codes so that encoder error does not leak into the decoder check
    m = net.forward_decode_by_code(rh.cuda(), rf.cuda()).cpu()
    ref = sho.forward_decode_by_code(shape_sd, rh, rf)
    assert float((m.sum(1) - 1).abs().max()) < 1e-5
    assert float((m - ref).abs().max()) < 1e-3, float((m - ref).abs().max())      # probabilities (measured 7.3e-4)
    agree = float((m.argmax(1) == ref.argmax(1)).float().mean())
    assert agree > 0.999, agree
    # VAE path returns (code, mean, std) like the reference (model.py:164-169)
    code, mean, std = net.forward_hair_encoder(hair.cuda())
    assert code.shape == (2, 16) and torch.equal(mean.cpu(), hc) and bool((std >= 0).all())
    with pytest.raises(_lib.ChbError):
        net.forward_face_encoder(face)  # host tensor
    bad = dict(shape_sd)
    bad.pop("face_decoder.out_layer.conv.bias")
    with pytest.raises(RuntimeError):
        ShapeGeneratorB200(max_batch=1).load_state_dict(bad)


@pytest.mark.gpu
def test_shape_other_batch_size_matches_oracle(shape_sd):
    """The plan follows the batch size (images per MMA tile, N tile and split-K factor of the 2x2 .. 8x8 layers and of
    the fully connected layers: csrc/shape.cu few_tiles_plan): a batch of 7 against the oracle, image by image."""
    from ctrlhair_b200.shape import ShapeGeneratorB200
    B = 7
    net = ShapeGeneratorB200(max_batch=B).load_state_dict(shape_sd)
    hair, face = synth.make_shape_inputs(B)
    hc = net.forward_hair_encoder(hair.cuda(), testing=True).cpu()
    fc = net.forward_face_encoder(face.cuda()).cpu()
    pick = [0, 3, 6]
    rh = sho.forward_hair_encoder(shape_sd, hair[pick])
    rf = sho.forward_face_encoder(shape_sd, face[pick])
    assert float((hc[pick] - rh).norm() / rh.norm()) < 1e-3
    assert float((fc[pick] - rf).norm() / rf.norm()) < 1e-3
    full_h, full_f = hc.clone(), fc.clone()
    full_h[pick], full_f[pick] = rh, rf          # decode from the oracle's codes where there are any
    m = net.forward_decode_by_code(full_h.cuda(), full_f.cuda()).cpu()
    ref = sho.forward_decode_by_code(shape_sd, rh, rf)
    assert float((m[pick] - ref).abs().max()) < 1e-3
    assert float((m[pick].argmax(1) == ref.argmax(1)).float().mean()) > 0.999


@pytest.mark.gpu
def test_shape_split_decoders_and_directly_change_hair_mask(shape_sd):
    """forward_hair_decoder / forward_face_decoder / forward_decoder (model.py:175-187) and the way
    Backend.directly_change_hair_mask composes them (ui/backend.py:409-420)."""
    from ctrlhair_b200.shape import ShapeGeneratorB200
    g = np.load(GOLD)
    net = ShapeGeneratorB200(max_batch=2).load_state_dict(shape_sd)
    rh, rf = torch.from_numpy(g["hair_code"]), torch.from_numpy(g["face_code"])
    hl = net.forward_hair_decoder(rh.cuda(), rf.cuda())
    fl = net.forward_face_decoder(rf.cuda())
    assert hl.shape == (2, 1, 256, 256) and fl.shape == (2, 18, 256, 256)
    ref_h = sho.mask_decoder(shape_sd, "hair_decoder", torch.cat([rf, rh], 1))
    ref_f = sho.mask_decoder(shape_sd, "face_decoder", rf)
    assert float((hl.cpu() - ref_h).norm() / ref_h.norm()) < 1e-3      # measured 6.7e-4 / 8.8e-4
    assert float((fl.cpu() - ref_f).norm() / ref_f.norm()) < 1e-3
    # the pieces compose to forward_decode_by_code bit for bit (same kernels, same logits)
    m = net.forward_decoder(hl, fl)
    assert torch.equal(m, net.forward_decode_by_code(rh.cuda(), rf.cuda()))
    # softmax on caller-made logits: exact same arithmetic as the oracle up to expf rounding
    hair_mask = (torch.rand((2, 1, 256, 256), generator=torch.Generator().manual_seed(3)) > 0.7).float()
    fl_c = fl.cpu()
    hair_logit = hair_mask * (fl_c.max() - fl_c.min() + 2) + fl_c.min() - 1          # ui/backend.py:417-418
    m2 = net.forward_decoder(hair_logit.cuda(), fl).cpu()
    ref2 = sho.forward_decoder(hair_logit, fl_c)
    assert float((m2 - ref2).abs().max()) < 1e-6
    assert torch.equal(m2.argmax(1), ref2.argmax(1))
    with pytest.raises(ValueError):
        net.forward_decoder(hl[:, :, :128], fl)


@pytest.mark.gpu
def test_encode_from_label_map_equals_one_hot_path(shape_sd):
    """chb_shape_encode_labels synthesises the one-hot planes inside the input gather: same codes as
    mask_label_to_one_hot + split_hair_face + the two encoders (ui/backend.py:81-86), label 255 = no class."""
    from ctrlhair_b200 import blend
    from ctrlhair_b200.shape import ShapeGeneratorB200
    B = 3
    shp = ShapeGeneratorB200(max_batch=B).load_state_dict(shape_sd)
    labels = synth.make_labels(B, 256, "blocky", seed=31)
    labels[0, :40, :40] = 255
    labels[1, 100:180, 60:200] = 13
    lab = labels.cuda()
    one_hot = blend.mask_label_to_one_hot(lab[:, None])
    hair, face = blend.split_hair_face(one_hot)
    want_h = shp.forward_hair_encoder(hair.contiguous(), testing=True)
    want_f = shp.forward_face_encoder(face.contiguous())
    got_h, got_f = shp.encode_labels(lab)
    # identical fp16 inputs to identical kernels; only the LayerNorm statistics (double atomics) can differ in the last bit
    assert float((got_h - want_h).abs().max()) <= 1e-5 * float(want_h.abs().max())
    assert float((got_f - want_f).abs().max()) <= 1e-5 * float(want_f.abs().max())


@pytest.mark.gpu
def test_decode_to_labels_equals_argmax_of_probabilities(shape_sd):
    """chb_shape_decode_labels == mask_one_hot_to_label(forward_decode_by_code(...)) (ui/backend.py:89-90)."""
    from ctrlhair_b200 import blend
    from ctrlhair_b200.shape import ShapeGeneratorB200
    B = 2
    shp = ShapeGeneratorB200(max_batch=B).load_state_dict(shape_sd)
    hair, face = synth.make_shape_inputs(B)
    hc, fc = shp.forward_hair_encoder(hair.cuda(), testing=True), shp.forward_face_encoder(face.cuda())
    want = blend.mask_one_hot_to_label(shp.forward_decode_by_code(hc, fc))
    got = shp.forward_decode_labels(hc, fc)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (B, 256, 256)
    assert float((got == want).float().mean()) > 0.9999     # two decoder runs: LN statistics use atomics
