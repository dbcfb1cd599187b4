"""Colour/texture branch nets on the fused MLP kernel, behind the reference's dict-in/dict-out surface.

Stand where `HairEditor.feature_generator / feature_encoder / feature_rgb_predictor` stand
(hair_editor.py:57-59; call sites ui/backend.py:96,103,167-169; solver.py:78-83 edit_infer) and load the same
state_dicts (`ckpt['Model_G']`, `ckpt['Model_D']`, `ckpt['Predictor']`, hair_editor.py:63-79) with strict keys.
Default dimensions are those of config 045 / predictor p004 (color_texture_branch/config.py:16-39,52-96,
predictor/predictor_config.py:30-43).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_LRELU, ACT_NONE, MlpLayer

BN_EPS = 1e-5


class _MlpChain:
    """A packed chain of dense layers living on the device + the C-ABI call."""

    def __init__(self, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.keep = []
        self.layers = []

    def add(self, weight, bias, pre_act=ACT_NONE, post_act=ACT_NONE, inj=None, zoff=0):
        """weight [out, in] (nn.Linear layout), stored transposed; inj = (U [nb, in], L [nb], mu [in])."""
        wt = weight.detach().float().t().contiguous().to(self.device)
        b = bias.detach().float().contiguous().to(self.device) if bias is not None else None
        L = MlpLayer()
        L.in_dim, L.out_dim = wt.shape[0], wt.shape[1]
        L.wt = wt.data_ptr()
        L.bias = b.data_ptr() if b is not None else None
        L.pre_act, L.post_act = pre_act, post_act
        self.keep += [wt, b]
        if inj is not None:
            U, Lv, mu = [t.detach().float().contiguous().to(self.device) for t in inj]
            L.inj_u, L.inj_l, L.inj_mu = U.data_ptr(), Lv.data_ptr(), mu.data_ptr()
            L.inj_nb, L.inj_zoff = U.shape[0], zoff
            self.keep += [U, Lv, mu]
        self.layers.append(L)

    def run(self, x, z=None):
        if not x.is_cuda:
            raise _lib.ChbError("colour/texture nets take CUDA tensors (there is no CPU path)")
        x = x.to(torch.float32).contiguous()
        B = x.shape[0]
        if x.dim() != 2 or x.shape[1] != self.layers[0].in_dim:
            raise _lib.ChbError("input has shape %s, expected [B, %d]" % (tuple(x.shape), self.layers[0].in_dim))
        out = torch.empty((B, self.layers[-1].out_dim), dtype=torch.float32, device=x.device)
        arr = (MlpLayer * len(self.layers))(*self.layers)
        zptr, zdim = None, 0
        if z is not None:
            z = z.to(device=x.device, dtype=torch.float32).contiguous()
            zptr, zdim = C.c_void_p(z.data_ptr()), z.shape[1]
        with torch.cuda.device(x.device):
            _lib.check(self.lib.chb_mlp_forward(arr, len(self.layers), C.c_void_p(x.data_ptr()), zptr, zdim,
                                                C.c_void_p(out.data_ptr()), B,
                                                C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
        return out


def _require_keys(sd, keys, what):
    missing = [k for k in keys if k not in sd]
    extra = [k for k in sd if k not in keys]
    if missing or extra:
        raise RuntimeError("Error(s) in loading state_dict for %s: missing %s unexpected %s" % (what, missing, extra))


class EigenGeneratorB200:
    """color_texture_branch/model_eigengan.py:34-83.  {'noise','noise_curliness','rgb_mean','pca_std'} -> {'code'}."""

    def __init__(self, device="cuda", hidden_layers=4, subspace_dim=2):
        self.device, self.nl, self.sd_dim = device, hidden_layers, subspace_dim
        self.chain = None

    def load_state_dict(self, sd, strict=True):
        keys = ["main_layer_in.weight", "main_layer_in.bias"]
        for i in range(self.nl):
            keys += ["main_layer_mid.%d.1.weight" % i, "main_layer_mid.%d.1.bias" % i]
        for i in range(self.nl):
            keys += ["subspaces.%d.U" % i, "subspaces.%d.L" % i, "subspaces.%d.mu" % i]
        if strict:
            _require_keys(sd, keys, "EigenGenerator")
        ch = _MlpChain(self.device)
        ch.add(sd["main_layer_in.weight"], sd["main_layer_in.bias"])
        for i in range(self.nl):
            # x = Linear_i(lrelu(x + subspace_i(z_i)))   (model_eigengan.py:78-81)
            ch.add(sd["main_layer_mid.%d.1.weight" % i], sd["main_layer_mid.%d.1.bias" % i], pre_act=ACT_LRELU,
                   inj=(sd["subspaces.%d.U" % i], sd["subspaces.%d.L" % i], sd["subspaces.%d.mu" % i]),
                   zoff=i * self.sd_dim)
        self.chain = ch
        return self

    def forward(self, data):
        # input order: curliness, rgb_mean, pca_std (model_eigengan.py:66-72)
        x = torch.cat([data["noise_curliness"], data["rgb_mean"], data["pca_std"]], dim=1)
        z = data["noise"].reshape(len(data["noise"]), self.nl * self.sd_dim)
        return {"code": self.chain.run(x, z)}

    __call__ = forward

    def eval(self):
        return self


class CodeEncoderB200:
    """The Discriminator used as encoder (color_texture_branch/model.py:86-127, config 045):
    {'code'} -> {'adv','noise','noise_curliness'}."""

    def __init__(self, device="cuda", hidden_layers=4, noise_dim=8, curliness_dim=1):
        self.device, self.nl, self.noise_dim, self.curl_dim = device, hidden_layers, noise_dim, curliness_dim
        self.chain = None

    def load_state_dict(self, sd, strict=True):
        keys = []
        for i in range(self.nl + 1):
            keys += ["net.%d.fc.weight" % i, "net.%d.fc.bias" % i]
        if strict:
            _require_keys(sd, keys, "Discriminator")
        ch = _MlpChain(self.device)
        for i in range(self.nl + 1):
            ch.add(sd["net.%d.fc.weight" % i], sd["net.%d.fc.bias" % i],
                   post_act=ACT_LRELU if i < self.nl else ACT_NONE)
        self.chain = ch
        return self

    def forward(self, data_in):
        out = self.chain.run(data_in["code"])
        p = 1 + self.noise_dim
        return {"adv": out[:, [0]], "noise": out[:, 1:p], "noise_curliness": out[:, p:p + self.curl_dim]}

    __call__ = forward

    def eval(self):
        return self


class PredictorB200:
    """color_texture_branch/predictor/predictor_model.py:14-41 in eval mode (BatchNorm1d running stats folded,
    dropout off): {'code'} -> {'rgb_mean','pca_std'}."""

    def __init__(self, device="cuda", hidden_layers=3, predict_dict=(("rgb_mean", 3), ("pca_std", 1))):
        self.device, self.nl, self.predict = device, hidden_layers, tuple(predict_dict)
        self.chain = None

    def load_state_dict(self, sd, strict=True):
        keys = []
        for i in range(self.nl):
            keys += ["net.%d.fc.weight" % i, "net.%d.fc.bias" % i] + \
                    ["net.%d.norm.%s" % (i, k) for k in ("weight", "bias", "running_mean", "running_var",
                                                         "num_batches_tracked")]
        keys += ["net.%d.fc.weight" % self.nl, "net.%d.fc.bias" % self.nl]
        if strict:
            _require_keys(sd, keys, "Predictor")
        ch = _MlpChain(self.device)
        for i in range(self.nl):
            # eval BatchNorm1d: y = (Wx + b - mean) * rstd * gamma + beta  -> folded into W, b
            s = sd["net.%d.norm.weight" % i].float() * torch.rsqrt(sd["net.%d.norm.running_var" % i].float() + BN_EPS)
            w = sd["net.%d.fc.weight" % i].float() * s[:, None]
            b = (sd["net.%d.fc.bias" % i].float() - sd["net.%d.norm.running_mean" % i].float()) * s + \
                sd["net.%d.norm.bias" % i].float()
            ch.add(w, b, post_act=ACT_LRELU)
        ch.add(sd["net.%d.fc.weight" % self.nl], sd["net.%d.fc.bias" % self.nl])
        self.chain = ch
        return self

    def forward(self, data_in):
        out = self.chain.run(data_in["code"])
        res, p = {}, 0
        for k, d in self.predict:
            res[k] = out[:, p:p + d]
            p += d
        return res

    __call__ = forward

    def eval(self):
        return self


def edit_infer(encoder, generator, hair_code, data):
    """color_texture_branch/solver.py:78-83: encode the code, override the given factors, decode."""
    inner = encoder({"code": hair_code})
    for k in data:
        inner[k] = data[k]
    return generator(inner)["code"]
