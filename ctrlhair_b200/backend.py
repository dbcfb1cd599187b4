"""Batched, device-resident mirror of the hot-path half of ui/backend.py::Backend.

The reference's Backend handles one image per call and bounces through the host between every stage
(`.cpu().numpy()` of the decoded mask at ui/backend.py:90,313, cv2 colour conversions at :98-125, 19 np.load calls per
gen_img at hair_editor.py:131-147).  BackendB200 keeps the same stages and names, for a batch of B images, with every
intermediate left on the GPU:

  parse_img(img_rgb, mask)        ui/backend.py:67-106   (get_mask / BiSeNet stays outside: the parsing is an input)
  set_input_img / set_target_img  ui/backend.py:127-145
  output(target_latent, feature)  ui/backend.py:147-175  (feature generator -> gen_img -> postprocess_blending)
  refresh_cur_mask                ui/backend.py:304-315
  transfer_latent_representation  ui/backend.py:266-302  ('color' / 'texture' / 'curliness'; 'shape' needs the ARAP warp
                                  of wrap_codes/, out of scope: pass the warped parsing with shape_from_mask)
  tensor_hsv_to_rgb / tensor_rgb_to_hsv / interpolate_hsv   ui/backend.py:108-125,323-332
  gen_img_batch(codes, parsing)   hair_editor.py:159-179 for B codes at once, median codes cached on the device
  change_curliness / change_color / change_shape / change_texture / continue_change_with_direction,
  get_*_be2fe, interpolate / interpolate_triple / interpolate_each_att, get_random_*, directly_change_hair_mask
                                  ui/backend.py:177-262,334-394,409-459: the latent edits of the UI, applied to every
                                  image of the batch (or to `index` only) without leaving the device
  DistTranslation                 util/color_from_hsv_to_gaussian.py:15-34 (slider value <-> HSV through the data set's
                                  sorted HSV table)

There is no CPU path: every network and every pre/post-processing step is a kernel of libctrlhair_b200.so.
"""
import copy
from bisect import bisect_left, bisect_right
from statistics import NormalDist

import numpy as np
import torch

from . import _lib, blend
from . import color_texture as ct
from .generator import SeanGeneratorB200
from .shape import ShapeGeneratorB200
from .zencoder import ZencoderB200

HAIR_IDX = 13


class LatentRepresentation:
    """ui/backend.py:31-37, every field batched [B, ...] on the device."""

    def __init__(self):
        self.color = None
        self.curliness = None
        self.shape = None
        self.texture = None
        self.face = None

    def clone(self):
        out = LatentRepresentation()
        for k in ("curliness", "shape", "texture", "face"):
            v = getattr(self, k)
            setattr(out, k, None if v is None else v.clone())
        out.color = None if self.color is None else {k: v.clone() for k, v in self.color.items()}
        return out


class DistTranslation:
    """util/color_from_hsv_to_gaussian.py:15-34 with the table passed in (the reference unpickles
    dataset_info_ctrlhair/hsv_stat_dict_ordered.pkl: uint8 [N,3], each column sorted).  scipy.stats.norm.cdf / ppf are
    the standard normal's; statistics.NormalDist gives the same values without the dependency."""

    def __init__(self, cols_hsv):
        self.cols_hsv = np.asarray(cols_hsv)
        self._n = NormalDist()

    def gaussian_to_val(self, dim, val):
        return self.cols_hsv[int(self._n.cdf(float(val)) * self.cols_hsv.shape[0])][dim]

    def val_to_gaussian(self, dim, val):
        col = self.cols_hsv[:, dim]
        left_v, right_v = bisect_left(col, val), bisect_right(col, val)
        q = (left_v + right_v) / 2 / self.cols_hsv.shape[0]
        if q <= 0.0 or q >= 1.0:            # scipy's ppf returns -inf / inf here
            return float("-inf") if q <= 0.0 else float("inf")
        return self._n.inv_cdf(q)


class BackendB200:
    def __init__(self, sean_sd, shape_sd, ct_sds, median_codes=None, max_batch=1, blending=True, img_size=256,
                 device=None, maximum_value_fe=2.5, hsv_table=None, shape_dirs=None, texture_dirs=None,
                 parsing_sd=None):
        g_sd, d_sd, p_sd = ct_sds
        self.netG = SeanGeneratorB200(crop=img_size, max_batch=max_batch, device=device).load_state_dict(sean_sd)
        self.zencoder = ZencoderB200(crop=img_size, max_batch=max_batch, device=device).load_state_dict(sean_sd)
        self.device = self.netG.device
        self.mask_generator = ShapeGeneratorB200(max_batch=max_batch, device=self.device).load_state_dict(shape_sd)
        self.feature_generator = ct.EigenGeneratorB200(device=self.device).load_state_dict(g_sd)
        self.feature_encoder = ct.CodeEncoderB200(device=self.device).load_state_dict(d_sd)
        self.feature_rgb_predictor = ct.PredictorB200(device=self.device).load_state_dict(p_sd)
        # face parsing (hair_editor.py:331-335 -> my_parsing_util.py:31-54): BiSeNet on the GPU when its checkpoint
        # (face_parsing_79999_iter.pth format) is given; otherwise callers pass the parsing themselves
        self.face_parser = None
        if parsing_sd is not None:
            from .bisenet import BiSeNetB200
            self.face_parser = BiSeNetB200(max_batch=max_batch, device=self.device,
                                           swap_labels=True).load_state_dict(parsing_sd)
        self.img_size = img_size
        self.max_batch = max_batch
        self.blending = blending
        self.seed = 0
        # hair_editor.py:131-147 reads 19 ACE.npy files on every gen_img; here they are uploaded once
        self.median = None if median_codes is None else torch.as_tensor(median_codes).float().to(self.device)
        self.maximum_value_fe = maximum_value_fe
        self.dist_translation = None if hsv_table is None else DistTranslation(hsv_table)
        self.shape_dirs = None if shape_dirs is None else [torch.as_tensor(d).float().to(self.device) for d in shape_dirs]
        self.texture_dirs = (None if texture_dirs is None
                             else [torch.as_tensor(d).float().to(self.device) for d in texture_dirs])
        self.input_img = self.input_mask = self.cur_mask = self.cur_latent = None
        self.input_sean_code = self.input_hair_feature = None
        self.target_img = self.target_mask = self.target_latent = self.target_hair_feature = None

    # ------------------------------------------------------------------ pre-processing (hair_editor.py:121-128)
    def preprocess_img(self, img_rgb):
        """uint8 [B,H,W,3] -> float [B,3,H,W] in [-1,1] (the cv2.resize to img_size is the caller's: sizes must match)."""
        t = torch.as_tensor(img_rgb).to(self.device)
        if t.shape[-3:-1] != (self.img_size, self.img_size):
            raise _lib.ChbError("images must already be %dx%d" % (self.img_size, self.img_size))
        return t.permute(0, 3, 1, 2).to(torch.float32) / 127.5 - 1.0

    def preprocess_mask(self, mask):
        t = torch.as_tensor(mask).to(self.device).to(torch.uint8)
        return t.reshape(-1, 1, self.img_size, self.img_size)

    # ------------------------------------------------------------------ colour space (ui/backend.py:108-125,323-332)
    def tensor_hsv_to_rgb(self, hsv):
        return blend.tensor_hsv_to_rgb(hsv, device=self.device)

    def tensor_rgb_to_hsv(self, rgb):
        return blend.tensor_rgb_to_hsv(rgb, device=self.device)

    def interpolate_hsv(self, hsv1, hsv2, alpha):
        rgb1 = self.tensor_hsv_to_rgb(hsv1)
        rgb2 = self.tensor_hsv_to_rgb(hsv2)
        rgb = rgb1 * (1 - alpha) + rgb2 * alpha
        return self.tensor_rgb_to_hsv(rgb)

    # ------------------------------------------------------------------ encode half (ui/backend.py:67-106)
    def shape_from_mask(self, mask_batch):
        """mask [B,1,S,S] uint8 -> (hair_code, face_code) (ui/backend.py:81-86)."""
        return self.mask_generator.encode_labels(mask_batch)   # one-hot + hair / face split happen inside the gather

    def get_code(self, img, mask_batch):
        """hair_editor.py:149-157: style codes [B,19,512] of the image under its parsing."""
        return self.zencoder(img.to(self.device, torch.float32).contiguous(), mask_batch[:, 0].contiguous())

    def get_mask(self, img_rgb):
        """HairEditor.get_mask (hair_editor.py:331-335) for a batch: uint8 [B,S,S,3] -> uint8 label maps [B,S,S] on the
        device (CelebAMask-HQ label order).  The PIL bilinear resize to the network's 512x512 (my_parsing_util.py:35) runs on
        the GPU as well, bit exact (bisenet.resize_bilinear_u8)."""
        if self.face_parser is None:
            raise _lib.ChbError("no face-parsing checkpoint was given (BackendB200(parsing_sd=...)); pass the mask")
        return self.face_parser.get_mask_device(img_rgb, self.img_size)

    def parse_img(self, img_rgb, mask=None, target_img=False):
        """Returns (img, out_mask, latent, mask, input_code, hair_feature) like ui/backend.py:67-106; `mask` is the
        parsing the reference gets from get_mask — computed here by the GPU face parser when it is None.  Everything is
        batched and on the device; out_mask is the uint8 label map [B,S,S] decoded from the shape codes (None for a
        target image)."""
        img_ts = torch.as_tensor(img_rgb).to(self.device)
        if mask is None:
            mask = self.get_mask(img_rgb)
        mask_batch = self.preprocess_mask(mask)
        lr = LatentRepresentation()
        out_mask = None
        if not target_img:
            lr.shape, lr.face = self.shape_from_mask(mask_batch)
            out_mask = self.mask_generator.forward_decode_labels(lr.shape, lr.face)   # softmax + argmax fused
        input_code = self.get_code(self.preprocess_img(img_ts), mask_batch)
        hair_feature = input_code[:, HAIR_IDX].contiguous()
        out_color = self.feature_rgb_predictor({"code": hair_feature})
        lr.color = {"hsv": self.tensor_rgb_to_hsv(out_color["rgb_mean"].contiguous()), "pca_std": out_color["pca_std"]}
        out_enc = self.feature_encoder({"code": hair_feature})
        lr.curliness = out_enc["noise_curliness"]
        lr.texture = out_enc["noise"]
        return img_ts, out_mask, lr, mask_batch[:, 0], input_code, hair_feature

    def set_input_img(self, img_rgb, mask=None):
        (self.input_img, self.cur_mask, self.cur_latent, self.input_mask, self.input_sean_code,
         self.input_hair_feature) = self.parse_img(img_rgb, mask)
        return self.input_img, self.cur_mask

    def set_target_img(self, img_rgb, mask=None):
        (self.target_img, _, self.target_latent, self.target_mask, _,
         self.target_hair_feature) = self.parse_img(img_rgb, mask, target_img=True)
        return self.target_img, self.target_mask

    # ------------------------------------------------------------------ decode half
    def refresh_cur_mask(self, target_latent=None):
        """ui/backend.py:304-315, label map stays on the device."""
        if target_latent is None:
            target_latent = self.cur_latent
        out_mask = self.mask_generator.forward_decode_labels(target_latent.shape, target_latent.face)
        self.cur_mask = out_mask
        return out_mask

    def gen_img_batch(self, codes, parsing, noise=None):
        """HairEditor.gen_img (hair_editor.py:159-179) for B (code, parsing) pairs: all-zero rows of a code are
        replaced by the median code of that class (:165-168).  codes [B,19,512], parsing uint8 [B,S,S] -> [B,3,S,S]."""
        codes = torch.as_tensor(codes).to(self.device, torch.float32)
        if self.median is None:
            raise _lib.ChbError("median style codes were not provided (hair_editor.py:134 reads them from disk)")
        empty = (codes == 0).all(dim=2, keepdim=True)
        codes = torch.where(empty, self.median[None].expand_as(codes), codes).contiguous()
        labels = torch.as_tensor(parsing).to(self.device).to(torch.uint8).reshape(-1, self.img_size, self.img_size)
        self.seed += 1
        return self.netG.forward_labels(labels.contiguous(), codes, noise=noise, seed=self.seed)

    def postprocess_blending(self, face_img, res_img, face_parsing, target_parsing, blending=True):
        return blend.postprocess_blending(face_img, res_img, face_parsing, target_parsing, blending=blending,
                                          device=self.device)

    def output(self, target_latent=None, feature=None, noise=None):
        """ui/backend.py:147-175 -> uint8 [B,S,S,3] on the device."""
        if target_latent is None:
            target_latent = self.cur_latent
            target_mask = self.cur_mask
        else:
            target_mask = self.refresh_cur_mask(target_latent)
        if "rgb_mean" in target_latent.color:
            target_color_rgb = target_latent.color["rgb_mean"]
        else:
            target_color_rgb = self.tensor_hsv_to_rgb(target_latent.color["hsv"])
        if feature is None:
            data = {"noise": target_latent.texture, "noise_curliness": target_latent.curliness,
                    "rgb_mean": target_color_rgb.to(torch.float32), "pca_std": target_latent.color["pca_std"]}
            feature = self.feature_generator(data)["code"]
        self.input_sean_code[:, HAIR_IDX] = feature
        edit_img = self.gen_img_batch(self.input_sean_code, target_mask, noise=noise)
        output_img, _ = self.postprocess_blending(self.input_img, edit_img, self.input_mask, target_mask,
                                                  blending=self.blending)
        return output_img

    def transfer_latent_representation(self, flag, refresh=True):
        """ui/backend.py:266-302 without the ARAP warp: for 'shape' set target_latent.shape / .face first
        (shape_from_mask on the warped target parsing)."""
        target_att = getattr(self.target_latent, flag)
        if target_att is None:
            raise _lib.ChbError("target latent has no '%s' (for 'shape' run shape_from_mask on the warped parsing)" % flag)
        if isinstance(target_att, torch.Tensor):
            setattr(self.cur_latent, flag, target_att.clone())
        else:
            cp = copy.copy(target_att)
            for k in cp:
                cp[k] = cp[k].clone()
            setattr(self.cur_latent, flag, cp)
        if flag == "shape" and refresh:
            self.refresh_cur_mask()
        if flag == "texture":
            self.transfer_latent_representation("curliness")

    # ------------------------------------------------------------------ latent edits (ui/backend.py:177-262,334-459)
    @staticmethod
    def _rows(t, index):
        return t if index is None else t[index:index + 1]

    def change_curliness(self, val, index=None):
        self._rows(self.cur_latent.curliness, index)[:] = val

    def change_color(self, val, idx, index=None):
        """idx 0 / 1 / 2 = hue / saturation / brightness through the data set's HSV distribution, idx 3 = variance in
        [-maximum_value_fe, maximum_value_fe] (ui/backend.py:196-209)."""
        if idx == 3:
            val = (val + self.maximum_value_fe) / 2 / self.maximum_value_fe
            self._rows(self.cur_latent.color["pca_std"], index)[:] = val * 100 + 20
        else:
            if self.dist_translation is None:
                raise _lib.ChbError("change_color needs the HSV table (hsv_table=...; the reference unpickles it)")
            self._rows(self.cur_latent.color["hsv"], index)[:, idx] = int(self.dist_translation.gaussian_to_val(idx, val))

    def continue_change_with_direction(self, att_name, direction, val, index=None):
        """att <- att + (val - <att, direction>) * direction, per image (ui/backend.py:450-459)."""
        direction = torch.as_tensor(direction).float().to(self.device)
        att = getattr(self.cur_latent, att_name)
        new = att + (val - att @ direction)[:, None] * direction[None]
        if index is not None:
            keep = torch.ones((att.shape[0], 1), device=att.device, dtype=torch.bool)
            keep[index] = False
            new = torch.where(keep, att, new)
        setattr(self.cur_latent, att_name, new)
        if att_name == "shape":
            self.refresh_cur_mask()

    def change_shape(self, val, idx, index=None):
        self.continue_change_with_direction("shape", self.shape_dirs[idx], val, index)

    def change_texture(self, val, idx, index=None):
        self.continue_change_with_direction("texture", self.texture_dirs[idx], val, index)

    def get_curliness_be2fe(self, index=0):
        return self.cur_latent.curliness[index]

    def get_color_be2fe(self, index=0):
        c_hsv = self.cur_latent.color["hsv"][index].cpu().numpy()
        color = [self.dist_translation.val_to_gaussian(k, c_hsv[k]) for k in range(3)]
        var_fe = ((self.cur_latent.color["pca_std"][index] - 20) / 100 * 2 * self.maximum_value_fe -
                  self.maximum_value_fe)
        return color[0], color[1], color[2], var_fe

    def get_shape_be2fe(self, index=0):
        return [torch.dot(self.cur_latent.shape[index], self.shape_dirs[i]) for i in range(len(self.shape_dirs))]

    def get_texture_be2fe(self, index=0):
        return [torch.dot(self.cur_latent.texture[index], self.texture_dirs[i]) for i in range(len(self.texture_dirs))]

    def interpolate(self, latent1, latent2, alpha):
        out = LatentRepresentation()
        for att in ("curliness", "shape", "texture"):
            setattr(out, att, getattr(latent1, att) * (1 - alpha) + getattr(latent2, att) * alpha)
        out.color = {"pca_std": latent1.color["pca_std"] * (1 - alpha) + latent2.color["pca_std"] * alpha,
                     "hsv": self.interpolate_hsv(latent1.color["hsv"], latent2.color["hsv"], alpha)}
        out.face = self.cur_latent.face
        return out

    def interpolate_triple(self, latent1, latent2, latent3, alpha1, alpha2, alpha3):
        latent12 = self.interpolate(latent1, latent2, alpha2 / (alpha1 + alpha2))
        return self.interpolate(latent12, latent3, alpha3)

    def interpolate_each_att(self, latent1, latent2, alpha, att_name):
        out = LatentRepresentation()
        for att in ("curliness", "shape", "texture"):
            setattr(out, att, getattr(self.cur_latent, att).clone())
        keep_color = {k: self.cur_latent.color[k].clone() for k in ("hsv", "pca_std")}
        if att_name == "shape":
            out.shape = latent1.shape * (1 - alpha) + latent2.shape * alpha
            out.color = keep_color
        elif att_name in ("curliness", "texture"):
            out.curliness = latent1.curliness * (1 - alpha) + latent2.curliness * alpha
            out.texture = latent1.texture * (1 - alpha) + latent2.texture * alpha
            out.color = keep_color
        else:
            out.color = {"pca_std": latent1.color["pca_std"] * (1 - alpha) + latent2.color["pca_std"] * alpha,
                         "hsv": self.interpolate_hsv(latent1.color["hsv"], latent2.color["hsv"], alpha)}
        out.face = self.cur_latent.face
        return out

    def _random_like(self, t, generator=None):
        return torch.randn(t.shape, generator=generator).to(self.device)    # common generate_noise: standard normal

    def get_random_texture(self, generator=None):
        self.cur_latent.texture = self._random_like(self.cur_latent.texture, generator)

    def get_random_shape(self, generator=None):
        self.cur_latent.shape = self._random_like(self.cur_latent.shape, generator)
        self.refresh_cur_mask()

    def get_random_curliness(self, generator=None):
        self.cur_latent.curliness = self._random_like(self.cur_latent.curliness, generator)

    def directly_change_hair_mask(self, hair_mask):
        """ui/backend.py:409-420: paste a hair region (label map [B,S,S], HAIR_IDX where hair) over the decoded face."""
        hair = torch.as_tensor(hair_mask).to(self.device) == HAIR_IDX
        face_logit = self.mask_generator.forward_face_decoder(self.cur_latent.face)
        hair_logit = hair.reshape(-1, 1, self.img_size, self.img_size).to(face_logit.dtype)
        hair_logit = hair_logit * (face_logit.max() - face_logit.min() + 2) + face_logit.min() - 1
        self.cur_mask = blend.mask_one_hot_to_label(self.mask_generator.forward_decoder(hair_logit, face_logit))
        return self.cur_mask
