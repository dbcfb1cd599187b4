"""The reference's call surface for the hot path: Pix2PixModel.forward(data, mode) and the HairEditor methods that
reach it, re-hosted on the B200 kernels.  Everything else of HairEditor / Backend (parsing, warping, blending)
stays where it is in the reference.

  Pix2PixModelB200.forward        sean_codes/models/pix2pix_model.py:39-74  (modes 'UI_mode', 'style_code')
  HairEditorB200.get_code         hair_editor.py:149-157
  HairEditorB200.gen_img          hair_editor.py:159-179
  HairEditorB200.generate_by_sean hair_editor.py:181-206
  HairEditorB200.load_average_feature  hair_editor.py:131-147
  HairEditorB200.gen_img_batch    batched gen_img for the validation / direction-finding callers
"""
import glob
import os

import numpy as np
import torch

from . import _lib
from .generator import SeanGeneratorB200
from .zencoder import ZencoderB200

HAIR_IDX = 13  # global_value_utils.py:49-52


class Pix2PixModelB200:
    def __init__(self, state_dict, crop=256, label_nc=19, max_batch=1, device=None):
        self.netG = SeanGeneratorB200(crop=crop, label_nc=label_nc, max_batch=max_batch, device=device)
        self.netG.load_state_dict(state_dict)
        self.zencoder = ZencoderB200(crop=crop, label_nc=label_nc, max_batch=max_batch, device=device)
        self.zencoder.load_state_dict(state_dict)
        self.netG.Zencoder = self.zencoder  # the reference reaches it as netG.Zencoder (pix2pix_model.py:71)
        self.device = self.netG.device
        self.label_nc = label_nc
        self.status = "test"
        self.seed = 0

    def eval(self):
        return self

    def modules(self):
        return [self, self.netG]

    def preprocess_input(self, data):
        """pix2pix_model.py:119-144: the label map goes to the device as class ids (the kernels build the one-hot
        pyramid themselves)."""
        label = data["label"]
        if label.dim() == 4:
            label = label[:, 0]
        return label.to(device=self.device, dtype=torch.uint8), data.get("image")

    def forward(self, data, mode):
        labels, image = self.preprocess_input(data)
        with torch.no_grad():
            if mode == "UI_mode":
                obj_dic = data["obj_dic"]
                codes = torch.stack([torch.as_tensor(obj_dic[str(j)]["ACE"]).float().reshape(-1)
                                     for j in range(self.label_nc)]).to(self.device)
                # The reference's UI_mode styles image 0 ONLY (`for i in range(1)`, normalization.py:124): images 1.. of a
                # batch would get a zero style map.  No caller relies on that, so a batch is rejected loudly instead of
                # being styled differently from the reference; batched callers use HairEditorB200.gen_img_batch.
                if labels.shape[0] != 1:
                    raise _lib.ChbError("UI_mode takes one image per call (the reference styles image 0 only, "
                                        "normalization.py:124); use gen_img_batch / forward_labels for batches")
                codes = codes[None].contiguous()
                self.seed += 1
                return self.netG.forward_labels(labels, codes, noise=data.get("noise"), seed=self.seed, graph=True)
            if mode == "style_code":
                return self.zencoder(image.to(self.device), labels)
        raise ValueError("|mode| is invalid")

    __call__ = forward


class HairEditorB200:
    """The slice of HairEditor that touches the generator and the style encoder."""

    def __init__(self, state_dict, median_codes=None, median_dir=None, img_size=256, device=None, max_batch=1):
        self.sean_model = Pix2PixModelB200(state_dict, crop=img_size, device=device, max_batch=max_batch)
        self.img_size = img_size
        self.device = self.sean_model.device
        if median_codes is None and median_dir is not None:
            median_codes = self.load_average_feature(median_dir)
        self.median = None if median_codes is None else torch.as_tensor(median_codes).float()

    @staticmethod
    def load_average_feature(folder):
        """<folder>/<class id>/ACE.npy, float32[512] each (hair_editor.py:131-147); classes without a file stay zero."""
        out = torch.zeros((19, 512), dtype=torch.float32)
        for i in range(19):
            files = sorted(glob.glob(os.path.join(folder, str(i), "*.npy")))
            for f in files:
                if os.path.splitext(os.path.basename(f))[0] == "ACE":
                    out[i] = torch.from_numpy(np.load(f)).float()
        return out

    def _obj_dic(self, code):
        if self.median is None:
            raise _lib.ChbError("median style codes were not provided (hair_editor.py:134 reads them from disk)")
        obj = {str(i): {"ACE": self.median[i].clone()} for i in range(19)}
        return obj

    def get_code(self, hair_img, hair_parsing):
        data = {"label": torch.as_tensor(hair_parsing, dtype=torch.float32), "instance": torch.tensor(0),
                "image": torch.as_tensor(hair_img, dtype=torch.float32), "path": ["temp/temp_npy"]}
        return self.sean_model(data, mode="style_code")

    def gen_img(self, code, parsing, noise=None):
        if not isinstance(code, torch.Tensor):
            code = torch.tensor(code)
        code = code.float().cpu()
        obj_dic = self._obj_dic(code)
        for idx in range(19):
            cur = code[0, idx]
            if not torch.all(cur == 0):
                obj_dic[str(idx)]["ACE"] = cur
        data = {"label": torch.as_tensor(parsing, dtype=torch.float32), "instance": torch.tensor(0),
                "image": torch.zeros((0, 3, self.img_size, self.img_size)), "obj_dic": obj_dic, "noise": noise}
        return self.sean_model(data, mode="UI_mode")[0]

    def gen_img_batch(self, codes, parsing, noise=None):
        """gen_img for B (code, parsing) pairs in one generator call (validation_in_train.py:87-288 and
        script_find_direction.py:55-74 render hundreds of images one gen_img at a time).  codes [B,19,512]; all-zero
        rows take the median code of their class (hair_editor.py:165-168), which is cached on the device instead of
        being re-read from 19 .npy files per call (:131-147).  parsing uint8 [B,S,S] -> images [B,3,S,S] (CUDA)."""
        if self.median is None:
            raise _lib.ChbError("median style codes were not provided (hair_editor.py:134 reads them from disk)")
        netG = self.sean_model.netG
        codes = torch.as_tensor(codes).to(self.device, torch.float32)
        B = codes.shape[0]
        if B > netG.max_batch:
            raise _lib.ChbError("batch %d exceeds the generator's max_batch %d" % (B, netG.max_batch))
        if getattr(self, "_median_dev", None) is None:
            self._median_dev = self.median.to(self.device)
        empty = (codes == 0).all(dim=2, keepdim=True)
        codes = torch.where(empty, self._median_dev[None].expand_as(codes), codes).contiguous()
        labels = torch.as_tensor(parsing).to(self.device).to(torch.uint8).reshape(B, self.img_size, self.img_size)
        self.sean_model.seed += 1
        return netG.forward_labels(labels.contiguous(), codes, noise=noise, seed=self.sean_model.seed)

    def generate_by_sean(self, face_img_code, hair_code, target_seg, noise=None):
        face_img_code = torch.as_tensor(face_img_code).float().cpu()
        obj_dic = self._obj_dic(face_img_code)
        for idx in range(19):
            cur = torch.as_tensor(hair_code).float().cpu() if idx == HAIR_IDX else face_img_code[idx]
            if not torch.all(face_img_code == 0):
                obj_dic[str(idx)]["ACE"] = cur
        data = {"label": torch.as_tensor(target_seg, dtype=torch.float32), "instance": torch.tensor(0),
                "obj_dic": obj_dic, "image": None, "noise": noise}
        return self.sean_model(data, mode="UI_mode")[0]
