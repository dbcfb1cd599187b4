"""Drop-in exercised, not asserted: the reference's REAL `Pix2PixModel` object with `netG` swapped for the B200 generator.

The unmodified reference tree is imported from baseline/_ref (staged by baseline/make_ref.py; it travels to the GPU
box) or, in the build container, from /root/reference.  The reference model is constructed the way HairEditor does it
(hair_editor.py:46-49: Pix2PixModel(opt) reads `<checkpoints_dir>/<name>/latest_net_G.pth` through util.load_network,
util/util.py:202-208), so the synthetic checkpoint goes through the reference's own loader too.

  * CPU (not gpu): the real model's 'UI_mode' output equals the oracle — pins the oracle through the true call surface.
  * GPU: `model.netG = SeanGeneratorB200(...)` (pix2pix_model.py:208-215 calls netG(seg, image, obj_dic=...)), then
    `model(data, 'UI_mode')` and `model(data, 'style_code')` against the unmodified model on the same inputs.
"""
import argparse
import contextlib
import os
import sys
import warnings

import pytest
import torch

from baseline import ref_runner
from ctrlhair_b200 import synth

needs_ref = pytest.mark.skipif(ref_runner.ref_root() is None, reason="reference tree not staged (baseline/make_ref.py)")
CROP = 64


def _opt(ckpt_dir, gpu):
    o = ref_runner.make_opt(crop=CROP)
    o.gpu_ids = [0] if gpu else []
    o.checkpoints_dir, o.name, o.which_epoch, o.continue_train = ckpt_dir, "synthetic", "latest", False
    return o


@contextlib.contextmanager
def _quiet():
    with warnings.catch_warnings(), open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
        warnings.simplefilter("ignore")
        yield


def _real_model(sd, tmp_path, gpu):
    d = os.path.join(str(tmp_path), "synthetic")
    os.makedirs(d, exist_ok=True)
    torch.save(sd, os.path.join(d, "latest_net_G.pth"))
    with ref_runner._on_path(ref_runner.ref_root(), cpu=not gpu), _quiet():
        from sean_codes.models.pix2pix_model import Pix2PixModel
        model = Pix2PixModel(_opt(str(tmp_path), gpu))
    model.eval()
    for m in model.modules():  # hair_editor.py:34-37
        if hasattr(m, "status"):
            m.status = "UI_mode"
    return model


def _data(labels, codes, image):
    # hair_editor.py:169-175 (gen_img) / :151-154 (get_code)
    return {"label": labels[:, None].float(), "instance": torch.tensor(0), "image": image,
            "obj_dic": {str(j): {"ACE": codes[0, j].clone()} for j in range(19)}, "path": ["temp/temp_npy"]}


@pytest.fixture(scope="module")
def small_sd(synthetic_sd):
    return synthetic_sd


@needs_ref
def test_real_pix2pix_model_ui_mode_equals_oracle_cpu(small_sd, tmp_path):
    from oracle import sean_oracle as so
    model = _real_model(small_sd, tmp_path, gpu=False)
    labels, codes, noise = synth.make_labels(1, CROP, "blocky"), synth.make_codes(1), synth.make_noise(1, CROP)
    empty = torch.zeros((0, 3, CROP, CROP))
    with ref_runner._on_path(ref_runner.ref_root(), cpu=True), ref_runner._injected_randn(noise, "cpu"), _quiet():
        got = model(_data(labels, codes, empty), "UI_mode")
    ref = so.generator_forward(small_sd, labels, codes, noise)
    assert got.shape == (1, 3, CROP, CROP)
    assert float((got - ref).abs().max() / ref.abs().max()) < 1e-4


@needs_ref
@pytest.mark.gpu
def test_netg_swapped_into_real_pix2pix_model(small_sd, tmp_path):
    from ctrlhair_b200 import _lib
    from ctrlhair_b200.generator import SeanGeneratorB200
    from ctrlhair_b200.zencoder import ZencoderB200
    torch.backends.cudnn.allow_tf32 = False  # the unmodified model is the fp32 yardstick here
    torch.backends.cuda.matmul.allow_tf32 = False
    model = _real_model(small_sd, tmp_path, gpu=True)
    labels, codes, noise = synth.make_labels(1, CROP, "blocky"), synth.make_codes(1), synth.make_noise(1, CROP)
    img = synth.make_image(1, CROP)
    empty = torch.zeros((0, 3, CROP, CROP))
    with ref_runner._on_path(ref_runner.ref_root(), cpu=False), _quiet():
        with ref_runner._injected_randn(noise, "cuda"):
            want_img = model(_data(labels, codes, empty), "UI_mode").cpu()
        want_codes = model(_data(labels, codes, img), "style_code").cpu()
        # ---- the swap a maintainer makes: only netG changes, Pix2PixModel stays the reference's
        netG = SeanGeneratorB200(crop=CROP, max_batch=1).load_state_dict(model.netG.state_dict())
        netG.Zencoder = ZencoderB200(crop=CROP, max_batch=1).load_state_dict(model.netG.state_dict())
        netG.fixed_noise = synth.flatten_noise(noise).cuda()
        model.netG = netG
        got_img = model(_data(labels, codes, empty), "UI_mode").cpu()
        got_codes = model(_data(labels, codes, img), "style_code").cpu()
        with pytest.raises(ValueError):
            model(_data(labels, codes, empty), "no_such_mode")
        with pytest.raises(_lib.ChbError):  # a UI_mode batch: the reference styles image 0 only -> rejected loudly
            netG(torch.zeros((2, 19, CROP, CROP)).cuda(), None, obj_dic=_data(labels, codes, empty)["obj_dic"])
    assert got_img.shape == want_img.shape == (1, 3, CROP, CROP) and got_img.is_contiguous()
    d = got_img - want_img
    assert float(d.norm() / want_img.norm()) < 1e-3 and float(d.abs().max() / want_img.abs().max()) < 1e-3
    assert got_codes.shape == want_codes.shape == (1, 19, 512)
    assert float((got_codes - want_codes).norm() / want_codes.norm()) < 1e-3
    absent = [j for j in range(19) if not bool((labels == j).any())]
    assert all(float(got_codes[0, j].abs().max()) == 0.0 for j in absent)
