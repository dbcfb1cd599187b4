"""BackendB200 (ctrlhair_b200/backend.py): the batched, device-resident Backend.set_input_img -> output chain
(ui/backend.py:67-106,127-175) including colour-space conversion, device label maps and Poisson blending, against the
same chain built from the CPU oracles.  Discontinuous intermediates (decoded label map, uint8 HSV colour) are taken
from the CUDA path when feeding the oracle's later stages, so that each stage is compared on identical inputs; their
own agreement is asserted separately."""
import numpy as np
import pytest
import torch

from ctrlhair_b200 import synth
from oracle import blend_oracle as bo
from oracle import ct_oracle as co
from oracle import sean_oracle as so
from oracle import shape_oracle as sho
from oracle import zencoder_oracle as zo

pytestmark = pytest.mark.gpu
HAIR = 13


def test_backend_set_input_output_chain(synthetic_sd):
    from ctrlhair_b200.backend import BackendB200
    B = 2
    shape_sd = synth.make_shape_state_dict()
    ct_sds = synth.make_ct_state_dicts()
    g_sd, d_sd, p_sd = ct_sds
    median = synth.make_codes(1, seed=4321)[0]
    cases = [synth.make_blend_case(256, 256, 40 + i) for i in range(B)]
    img_u8 = np.stack([c[0] for c in cases])                                  # input faces, uint8 HWC
    labels = torch.from_numpy(np.stack([c[2] for c in cases]))                # their parsings (hair, skin, bg, mouth)
    noise = synth.make_noise(B, 256)

    be = BackendB200(synthetic_sd, shape_sd, ct_sds, median_codes=median, max_batch=B, blending=True)
    img_ts, cur_mask, lat, in_mask, in_code, hair_feat = be.parse_img(img_u8, labels)
    be.set_input_img(img_u8, labels)
    assert torch.equal(be.cur_mask, cur_mask) and torch.equal(be.input_mask, labels.cuda())

    # ---- encode half vs the oracles
    oh = so.one_hot(labels)
    hair, face = oh[:, [HAIR]], torch.cat([oh[:, :HAIR], oh[:, HAIR + 1:]], 1)
    r_hc, r_fc = sho.forward_hair_encoder(shape_sd, hair), sho.forward_face_encoder(shape_sd, face)
    assert float((lat.shape.cpu() - r_hc).norm() / r_hc.norm()) < 3e-3
    assert float((lat.face.cpu() - r_fc).norm() / r_fc.norm()) < 3e-3
    r_mask = sho.forward_decode_by_code(shape_sd, r_hc, r_fc)
    r_lab = torch.from_numpy(bo.mask_one_hot_to_label(r_mask.numpy()).astype(np.uint8))
    assert float((cur_mask.cpu() == r_lab).float().mean()) > 0.995
    img_f = torch.from_numpy(img_u8).permute(0, 3, 1, 2).float() / 127.5 - 1.0      # hair_editor.py:121-123
    r_codes = zo.zencoder_forward(synthetic_sd, img_f, labels)
    assert float((in_code.cpu() - r_codes).norm() / r_codes.norm()) < 2e-3
    assert bool((r_codes == 0).all(2).any())       # some class is absent: its row must fall back to the median code
    r_pred = co.predictor(p_sd, {"code": r_codes[:, HAIR]})
    pred = be.feature_rgb_predictor({"code": hair_feat})
    assert float((pred["rgb_mean"].cpu() - r_pred["rgb_mean"]).abs().max()) < 5e-3 * float(r_pred["rgb_mean"].abs().max())
    hsv = lat.color["hsv"].cpu().numpy()
    assert hsv.dtype == np.uint8 and hsv.shape == (B, 3)
    assert np.array_equal(hsv, bo.rgb_to_hsv_u8(bo.float_to_u8_trunc(pred["rgb_mean"].cpu().numpy())))  # ui/backend.py:98-99
    r_enc = co.discriminator(d_sd, {"code": r_codes[:, HAIR]})
    assert float((lat.texture.cpu() - r_enc["noise"]).abs().max()) < 5e-3 * float(r_enc["noise"].abs().max())

    # ---- an edit of the colour, then the decode half
    be.cur_latent.color["hsv"][:, 0] = (be.cur_latent.color["hsv"][:, 0].int() + 40).remainder(180).to(torch.uint8)
    out = be.output(noise=synth.flatten_noise(noise).cuda())
    assert out.dtype == torch.uint8 and tuple(out.shape) == (B, 256, 256, 3)

    rgb = bo.hsv_to_rgb_u8(be.cur_latent.color["hsv"].cpu().numpy())                # ui/backend.py:108-115
    data = {"noise": lat.texture.cpu(), "noise_curliness": lat.curliness.cpu(),
            "rgb_mean": torch.from_numpy(rgb).float(), "pca_std": lat.color["pca_std"].cpu()}
    r_feat = co.eigen_generator(g_sd, data)["code"]
    r_in = in_code.cpu().clone()
    r_in[:, HAIR] = r_feat                                                         # ui/backend.py:170
    assert float((be.input_sean_code[:, HAIR].cpu() - r_feat).norm() / r_feat.norm()) < 5e-3
    empty = (r_in == 0).all(2, keepdim=True)
    r_in = torch.where(empty, median[None].expand_as(r_in), r_in)                  # hair_editor.py:165-168
    r_img = so.generator_forward(synthetic_sd, cur_mask.cpu(), r_in, noise)
    edit = be.gen_img_batch(be.input_sean_code, cur_mask, noise=synth.flatten_noise(noise).cuda())
    assert float((edit.cpu() - r_img).norm() / r_img.norm()) < 2e-3

    # blending stage alone, on the CUDA path's own generated image: the solved bytes agree with spsolve
    for i in range(B):
        want, want_mask = bo.postprocess_blending(img_u8[i], edit[i].cpu().numpy(), labels[i].numpy(),
                                                  cur_mask[i].cpu().numpy())
        d = np.abs(out[i].cpu().numpy().astype(int) - want.astype(int))
        assert d.max() <= 1
        solved = bo.unknown_set(1 - want_mask[..., 0])
        assert int((d[solved] != 0).sum()) <= max(3, int(1e-4 * d[solved].size))
        # whole chain against the all-oracle image: fp16 generator error (1e-3 relative) moves a few uint8 levels
        want_all, _ = bo.postprocess_blending(img_u8[i], r_img[i].numpy(), labels[i].numpy(), cur_mask[i].cpu().numpy())
        d = np.abs(out[i].cpu().numpy().astype(int) - want_all.astype(int))
        assert d.max() <= 4 and d.mean() < 0.6, (d.max(), d.mean())

    # blending off: the plain uint8 image (hair_editor.py:307-308)
    be.blending = False
    plain = be.output(noise=synth.flatten_noise(noise).cuda())
    assert np.array_equal(plain.cpu().numpy(), np.stack([bo.tensor_to_cv2_u8(e) for e in edit.cpu().numpy()]))

    # transfer of colour / texture latents from a target image (ui/backend.py:266-302)
    be.set_target_img(np.stack([c[1] for c in cases]), labels)
    be.transfer_latent_representation("texture")
    assert torch.equal(be.cur_latent.texture, be.target_latent.texture)
    assert torch.equal(be.cur_latent.curliness, be.target_latent.curliness)
    h1, h2 = be.cur_latent.color["hsv"], be.target_latent.color["hsv"]
    mix = be.interpolate_hsv(h1, h2, 0.25).cpu().numpy()
    r1, r2 = bo.hsv_to_rgb_u8(h1.cpu().numpy()).astype(np.float32), bo.hsv_to_rgb_u8(h2.cpu().numpy()).astype(np.float32)
    assert np.array_equal(mix, bo.rgb_to_hsv_u8(bo.float_to_u8_trunc(r1 * 0.75 + r2 * 0.25)))


def test_backend_latent_edits(synthetic_sd):
    """The UI's latent edits (ui/backend.py:177-262,334-459) on a batch, against the reference formulas."""
    import scipy.stats as st
    from bisect import bisect_left, bisect_right
    from ctrlhair_b200.backend import BackendB200, DistTranslation
    B = 2
    shape_sd = synth.make_shape_state_dict()
    g = np.random.default_rng(3)
    table = np.sort(g.integers(0, 256, (5000, 3)), axis=0).astype(np.uint8)     # hsv_stat_dict_ordered.pkl stand-in
    table[:, 0] = np.sort(g.integers(0, 180, 5000)).astype(np.uint8)
    tg = torch.Generator().manual_seed(9)
    sdirs = [torch.nn.functional.normalize(torch.randn(16, generator=tg), dim=0) for _ in range(4)]
    tdirs = [torch.nn.functional.normalize(torch.randn(8, generator=tg), dim=0) for _ in range(2)]
    be = BackendB200(synthetic_sd, shape_sd, synth.make_ct_state_dicts(), median_codes=synth.make_codes(1, seed=4321)[0],
                     max_batch=B, hsv_table=table, shape_dirs=sdirs, texture_dirs=tdirs)
    cases = [synth.make_blend_case(256, 256, 60 + i) for i in range(B)]
    labels = torch.from_numpy(np.stack([c[2] for c in cases]))
    be.set_input_img(np.stack([c[0] for c in cases]), labels)
    be.set_target_img(np.stack([c[1] for c in cases]), labels)

    # util/color_from_hsv_to_gaussian.py:22-34 with scipy, as the reference computes it
    dt = DistTranslation(table)
    for dim, v in [(0, -1.3), (1, 0.0), (2, 2.2)]:
        assert dt.gaussian_to_val(dim, v) == table[int(st.norm.cdf(v) * table.shape[0])][dim]
    for dim, val in [(0, 90), (1, 17), (2, 250)]:
        want = st.norm.ppf((bisect_left(table[:, dim], val) + bisect_right(table[:, dim], val)) / 2 / table.shape[0])
        assert abs(dt.val_to_gaussian(dim, val) - want) < 1e-9

    # colour sliders
    be.change_color(0.7, 1)
    assert bool((be.cur_latent.color["hsv"][:, 1] == int(dt.gaussian_to_val(1, 0.7))).all())
    be.change_color(1.0, 3, index=1)
    assert abs(float(be.cur_latent.color["pca_std"][1]) - ((1.0 + 2.5) / 5.0 * 100 + 20)) < 1e-4
    c0, c1, c2, var_fe = be.get_color_be2fe(index=1)
    assert abs(float(var_fe) - 1.0) < 1e-5 and abs(c1 - dt.val_to_gaussian(1, int(be.cur_latent.color["hsv"][1, 1]))) < 1e-12

    # projection edits: <att, direction> becomes val, the orthogonal part is untouched; 'shape' refreshes the mask
    before = be.cur_latent.texture.clone()
    be.change_texture(0.8, 1)
    proj = be.cur_latent.texture @ tdirs[1].cuda()
    assert float((proj - 0.8).abs().max()) < 1e-5
    want = before + (0.8 - before @ tdirs[1].cuda())[:, None] * tdirs[1].cuda()[None]
    assert torch.allclose(be.cur_latent.texture, want, atol=1e-6)
    assert abs(float(be.get_texture_be2fe(0)[1]) - 0.8) < 1e-5
    mask_before = be.cur_mask.clone()
    shape_before = be.cur_latent.shape.clone()
    be.change_shape(3.0, 0, index=0)
    assert torch.equal(be.cur_latent.shape[1], shape_before[1])               # only image 0 edited
    assert abs(float(be.get_shape_be2fe(0)[0]) - 3.0) < 1e-4
    assert not torch.equal(be.cur_mask[0], mask_before[0])
    assert float((be.cur_mask[1] == mask_before[1]).float().mean()) > 0.999   # (LN statistics use atomics)
    be.change_curliness(-0.5)
    assert float((be.cur_latent.curliness + 0.5).abs().max()) == 0.0

    # interpolation (ui/backend.py:334-394)
    l1, l2 = be.cur_latent.clone(), be.target_latent.clone()
    l2.shape, l2.face = l1.shape + 1.0, l1.face
    mix = be.interpolate(l1, l2, 0.25)
    assert torch.allclose(mix.texture, l1.texture * 0.75 + l2.texture * 0.25)
    r1 = bo.hsv_to_rgb_u8(l1.color["hsv"].cpu().numpy()).astype(np.float32)
    r2 = bo.hsv_to_rgb_u8(l2.color["hsv"].cpu().numpy()).astype(np.float32)
    assert np.array_equal(mix.color["hsv"].cpu().numpy(), bo.rgb_to_hsv_u8(bo.float_to_u8_trunc(r1 * 0.75 + r2 * 0.25)))
    tri = be.interpolate_triple(l1, l2, l1, 1.0, 1.0, 0.5)
    assert torch.allclose(tri.curliness, (l1.curliness * 0.5 + l2.curliness * 0.5) * 0.5 + l1.curliness * 0.5)
    only = be.interpolate_each_att(l1, l2, 0.5, "shape")
    assert torch.allclose(only.shape, l1.shape + 0.5) and torch.equal(only.texture, be.cur_latent.texture)
    assert torch.equal(only.color["hsv"], be.cur_latent.color["hsv"])

    # pasting a hair region over the decoded face (ui/backend.py:409-420)
    hair = np.zeros((B, 256, 256), np.uint8)
    hair[:, 40:120, 60:200] = HAIR
    lab = be.directly_change_hair_mask(hair).cpu().numpy()
    assert (lab[:, 40:120, 60:200] == HAIR).all() and not (lab[:, 150:, :] == HAIR).any()

    # random re-draws keep the shapes and refresh the mask
    be.get_random_shape(generator=torch.Generator().manual_seed(1))
    assert tuple(be.cur_latent.shape.shape) == (B, 16) and tuple(be.cur_mask.shape) == (B, 256, 256)
    be.get_random_texture()
    be.get_random_curliness()
    out = be.output()
    assert out.dtype == torch.uint8 and tuple(out.shape) == (B, 256, 256, 3)
