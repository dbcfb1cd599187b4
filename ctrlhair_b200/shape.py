"""Host-side mirror of the reference shape generator (`HairEditor.mask_generator`, shape_branch/model.py:146-199)."""
import ctypes as C

import torch

from . import _lib

KY_MAP = {0: (-1, 1), 1: (0, 0), 2: (0, 1), 3: (1, 0)}  # 4x4/s2 tap -> (block offset, parity) in space-to-depth


def pack_s2d_conv(w, cin_pad_total):
    """conv4x4 stride 2 pad 1 weight [co, cin, 4, 4] -> conv3x3 over the space-to-depth map:
    [co, 9 * cin_pad_total] with k = tap * cin_pad_total + (py*2+px) * cin + c."""
    co, cin = w.shape[0], w.shape[1]
    out = torch.zeros((co, 3, 3, cin_pad_total), dtype=w.dtype)
    for ky in range(4):
        dy, py = KY_MAP[ky]
        for kx in range(4):
            dx, px = KY_MAP[kx]
            par = py * 2 + px
            out[:, dy + 1, dx + 1, par * cin:(par + 1) * cin] = w[:, :, ky, kx]
    return out.reshape(co, 9 * cin_pad_total)


def _k_major(w):
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def pack_shape(sd, weight_dtype=torch.float16):
    out = {}
    for net, cm, pad0 in (("hair", 1, 192), ("face", 18, 256)):
        p = net + "_encoder."
        cin_tot = pad0
        for i in range(7):
            q = "%s_encoder.layers.%d." % (net, i)
            w = sd[q + "conv.weight"].float()
            out[p + "%d.w" % i] = pack_s2d_conv(w, cin_tot).to(weight_dtype)
            out[p + "%d.b" % i] = sd[q + "conv.bias"].float()
            out[p + "%d.gamma" % i] = sd[q + "norm.gamma"].float()
            out[p + "%d.beta" % i] = sd[q + "norm.beta"].float()
            cin_tot = 4 * w.shape[0]
        # fc: reference flattens NCHW [2048,2,2]; here the feature is NHWC [2,2,2048]
        ws = [sd["%s_encoder.out_layer.fc.weight" % net].float()]
        bs = [sd["%s_encoder.out_layer.fc.bias" % net].float()]
        if net == "hair":
            ws.append(sd["hair_encoder.std_out_layer.fc.weight"].float())
            bs.append(sd["hair_encoder.std_out_layer.fc.bias"].float())
        w = torch.cat(ws, 0)
        w = w.reshape(w.shape[0], 2048, 2, 2).permute(0, 2, 3, 1).reshape(w.shape[0], 8192)
        out[p + "fc.w"] = w.contiguous().to(weight_dtype)
        out[p + "fc.b"] = torch.cat(bs, 0)
    for net, kpad, rows in (("hair", 1088, 32), ("face", 1024, 32)):
        p = net + "_decoder."
        w = sd["%s_decoder.in_layer.fc.weight" % net].float()       # [8192, in]
        b = sd["%s_decoder.in_layer.fc.bias" % net].float()
        wp = torch.zeros((8192, kpad))
        wp[:, :w.shape[1]] = w
        perm = wp.reshape(2048, 2, 2, kpad).permute(1, 2, 0, 3).reshape(8192, kpad)   # rows (c,y,x) -> (y,x,c)
        out[p + "fc.w"] = perm.contiguous().to(weight_dtype)
        out[p + "fc.b"] = b.reshape(2048, 2, 2).permute(1, 2, 0).reshape(8192).contiguous()
        for i in range(7):
            q = "%s_decoder.layers.%d." % (net, 2 * i + 1)
            out[p + "%d.w" % i] = _k_major(sd[q + "conv.weight"].float()).to(weight_dtype)
            out[p + "%d.b" % i] = sd[q + "conv.bias"].float()
            out[p + "%d.gamma" % i] = sd[q + "norm.gamma"].float()
            out[p + "%d.beta" % i] = sd[q + "norm.beta"].float()
        w = sd["%s_decoder.out_layer.conv.weight" % net].float()
        wo = torch.zeros((rows, 9 * 32))
        wo[:w.shape[0]] = _k_major(w)
        bo = torch.zeros(rows)
        bo[:w.shape[0]] = sd["%s_decoder.out_layer.conv.bias" % net].float()
        out[p + "out.w"] = wo.to(weight_dtype)
        out[p + "out.b"] = bo
    return out


class ShapeGeneratorB200:
    def __init__(self, max_batch=1, device=None):
        if not torch.cuda.is_available():
            raise _lib.ChbError("ShapeGeneratorB200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.max_batch = max_batch
        cfg = _lib.ShapeConfig(256, max_batch)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_check_device())
            _lib.check(self.lib.chb_shape_create(C.byref(cfg), C.byref(h)))
        self.handle, self.blob, self.workspace = h, None, None

    def __del__(self):
        # (at interpreter shutdown torch.nn may already be torn down: bypass nn.Module.__setattr__, never raise)
        h = self.__dict__.get("handle")
        if h:
            self.__dict__["handle"] = None
            try:
                self.lib.chb_shape_destroy(h)
            except Exception:
                pass

    def eval(self):
        return self

    def _layout(self):
        n = self.lib.chb_shape_num_tensors(self.handle)
        name = C.create_string_buffer(96)
        off, nb, dt = C.c_int64(), C.c_int64(), C.c_int()
        lay = {}
        for i in range(n):
            _lib.check(self.lib.chb_shape_tensor_info(self.handle, i, name, 96, C.byref(off), C.byref(nb), C.byref(dt)))
            lay[name.value.decode()] = (off.value, nb.value, dt.value)
        return lay

    def load_state_dict(self, sd, strict=True):
        from .synth import shape_shapes
        if strict:
            want = set(shape_shapes())
            if set(sd) != want:
                raise RuntimeError("Error(s) in loading state_dict for Generator: missing %s unexpected %s" %
                                   (sorted(want - set(sd)), sorted(set(sd) - want)))
        packed = pack_shape(sd)
        lay = self._layout()
        if set(lay) != set(packed):
            raise _lib.ChbError("packer/library layout mismatch: %s" % sorted(set(lay) ^ set(packed)))
        blob = torch.zeros(self.lib.chb_shape_blob_bytes(self.handle), dtype=torch.uint8)
        for k, (off, nb, dt) in lay.items():
            t = packed[k].contiguous()
            want_dt = torch.float16 if dt == _lib.F16 else torch.float32
            if t.dtype != want_dt or t.numel() * t.element_size() != nb:
                raise _lib.ChbError("packed tensor %s: %s x %d bytes, library wants %d" %
                                    (k, t.dtype, t.numel() * t.element_size(), nb))
            blob[off:off + nb] = t.view(torch.uint8).reshape(-1)
        self.blob = blob.to(self.device)
        self.workspace = torch.empty(self.lib.chb_shape_workspace_bytes(self.handle) + 1024, dtype=torch.uint8,
                                     device=self.device)
        ws = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_shape_bind(self.handle, C.c_void_p(self.blob.data_ptr()), C.c_void_p(ws)))
        return self

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _encode(self, net, mask, width):
        if not mask.is_cuda:
            raise _lib.ChbError("shape nets take CUDA tensors (there is no CPU path)")
        mask = mask.to(torch.float32).contiguous()
        B = mask.shape[0]
        out = torch.empty((B, width), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_shape_encode(self.handle, net, C.c_void_p(mask.data_ptr()),
                                                 C.c_void_p(out.data_ptr()), B, self._stream()))
        return out

    def encode_labels(self, labels):
        """Both encoders straight from a label map uint8 [B,S,S] (or [B,1,S,S]): what Backend.parse_img computes with
        mask_label_to_one_hot + split_hair_face + forward_hair_encoder(testing=True) + forward_face_encoder
        (ui/backend.py:81-86), without materialising the one-hot tensors.  Returns (hair_code [B,16], face_code)."""
        if not labels.is_cuda:
            raise _lib.ChbError("shape nets take CUDA tensors (there is no CPU path)")
        labels = labels.to(torch.uint8).reshape(-1, 256, 256).contiguous()
        B = labels.shape[0]
        outs = []
        for net, width in ((0, 32), (1, 1024)):
            out = torch.empty((B, width), dtype=torch.float32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(self.lib.chb_shape_encode_labels(self.handle, net, C.c_void_p(labels.data_ptr()),
                                                            C.c_void_p(out.data_ptr()), B, self._stream()))
            outs.append(out)
        return outs[0][:, :16], outs[1]

    def forward_hair_encoder(self, hair, testing=False):
        out = self._encode(0, hair, 32)
        mean, std = out[:, :16], out[:, 16:].abs()
        if testing:
            return mean
        code = torch.randn_like(mean) * std + mean   # vae_resampling (model.py:110-113)
        return code, mean, std

    def forward_face_encoder(self, face):
        return self._encode(1, face, 1024)

    def forward_decode_by_code(self, hair_code, face_code):
        B = hair_code.shape[0]
        hair_code = hair_code.to(device=self.device, dtype=torch.float32).contiguous()
        face_code = face_code.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty((B, 19, 256, 256), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_shape_decode(self.handle, C.c_void_p(hair_code.data_ptr()),
                                                 C.c_void_p(face_code.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                                 self._stream()))
        return out

    def forward_decode_labels(self, hair_code, face_code):
        """forward_decode_by_code + shape_util.mask_one_hot_to_label in one call (ui/backend.py:89-90,312-313): uint8
        label map [B,256,256], the [B,19,256,256] probabilities are never written."""
        B = hair_code.shape[0]
        hair_code = hair_code.to(device=self.device, dtype=torch.float32).contiguous()
        face_code = face_code.to(device=self.device, dtype=torch.float32).contiguous()
        out = torch.empty((B, 256, 256), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_shape_decode_labels(self.handle, C.c_void_p(hair_code.data_ptr()),
                                                        C.c_void_p(face_code.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                                        self._stream()))
        return out

    def _decode_logits(self, net, hair_code, face_code, channels):
        B = face_code.shape[0]
        face_code = face_code.to(device=self.device, dtype=torch.float32).contiguous()
        hp = None
        if hair_code is not None:
            hair_code = hair_code.to(device=self.device, dtype=torch.float32).contiguous()
            hp = C.c_void_p(hair_code.data_ptr())
        out = torch.empty((B, channels, 256, 256), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_shape_decode_logits(self.handle, net, hp, C.c_void_p(face_code.data_ptr()),
                                                        C.c_void_p(out.data_ptr()), B, self._stream()))
        return out

    def forward_hair_decoder(self, hair_code, face_code):
        """shape_branch/model.py:175-178: hair logit [B,1,256,256] from cat([face_code, hair_code])."""
        return self._decode_logits(0, hair_code, face_code, 1)

    def forward_face_decoder(self, face_code):
        """shape_branch/model.py:180-182: face logits [B,18,256,256] (ui/backend.py:416)."""
        return self._decode_logits(1, None, face_code, 18)

    def forward_decoder(self, hair_logit, face_logit):
        """shape_branch/model.py:184-187: softmax over [face[:13], hair, face[13:]] (ui/backend.py:419)."""
        B = face_logit.shape[0]
        if not (hair_logit.is_cuda and face_logit.is_cuda):
            raise _lib.ChbError("shape nets take CUDA tensors (there is no CPU path)")
        if tuple(hair_logit.shape) != (B, 1, 256, 256) or tuple(face_logit.shape) != (B, 18, 256, 256):
            raise ValueError("forward_decoder wants hair [B,1,256,256] and face [B,18,256,256] logits")
        hair_logit = hair_logit.to(torch.float32).contiguous()
        face_logit = face_logit.to(torch.float32).contiguous()
        out = torch.empty((B, 19, 256, 256), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_shape_softmax(self.handle, C.c_void_p(hair_logit.data_ptr()),
                                                  C.c_void_p(face_logit.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                                  self._stream()))
        return out

    def forward_edit_directly_in_test(self, hair, face):
        return self.forward_decode_by_code(self.forward_hair_encoder(hair, testing=True), self.forward_face_encoder(face))
