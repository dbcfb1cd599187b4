"""Host-side mirror of the reference style encoder (`netG.Zencoder`, architecture.py:154-207) on the C ABI."""
import ctypes as C

import torch

from . import _lib


def pack_zencoder(sd, weight_dtype=torch.float16):
    """Reference netG state_dict -> {blob tensor name: CPU tensor}.  Conv weights [rows][tap*C + c];
    the ConvTranspose2d (k3 s2 p1 op1) becomes a stride-1 conv over the zero-inserted map with the kernel flipped
    and its in/out axes swapped."""
    p = "Zencoder.model."
    out = {}
    # L1 stays fp32: [32][(ky, kx, ci)]
    out["l1.w"] = sd[p + "1.weight"].float().permute(0, 2, 3, 1).reshape(32, 27).contiguous()
    out["l1.b"] = sd[p + "1.bias"].float()

    def km(w):
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

    out["l2.w"] = km(sd[p + "4.weight"].float()).to(weight_dtype)
    out["l2.b"] = sd[p + "4.bias"].float()
    out["l3.w"] = km(sd[p + "7.weight"].float()).to(weight_dtype)
    out["l3.b"] = sd[p + "7.bias"].float()
    wt = sd[p + "10.weight"].float()                       # [in=128, out=256, 3, 3]
    out["l4.w"] = km(wt.flip(2, 3).permute(1, 0, 2, 3)).to(weight_dtype)
    out["l4.b"] = sd[p + "10.bias"].float()
    out["l5.w"] = km(sd[p + "14.weight"].float()).to(weight_dtype)
    out["l5.b"] = sd[p + "14.bias"].float()
    return out


class ZencoderB200(torch.nn.Module):
    """nn.Module for the same reason as SeanGeneratorB200: the reference reaches it as the child module
    `netG.Zencoder` (pix2pix_model.py:71)."""

    def __init__(self, crop=256, label_nc=19, max_batch=1, device=None):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.ChbError("ZencoderB200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.crop, self.label_nc, self.max_batch = crop, label_nc, max_batch
        cfg = _lib.ZencConfig(crop, label_nc, max_batch)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_check_device())
            _lib.check(self.lib.chb_zencoder_create(C.byref(cfg), C.byref(h)))
        self.handle, self.blob, self.workspace = h, None, None

    def __del__(self):
        # (at interpreter shutdown torch.nn may already be torn down: bypass nn.Module.__setattr__, never raise)
        h = self.__dict__.get("handle")
        if h:
            self.__dict__["handle"] = None
            try:
                self.lib.chb_zencoder_destroy(h)
            except Exception:
                pass

    def _layout(self):
        n = self.lib.chb_zencoder_num_tensors(self.handle)
        name = C.create_string_buffer(64)
        off, nb, dt = C.c_int64(), C.c_int64(), C.c_int()
        lay = {}
        for i in range(n):
            _lib.check(self.lib.chb_zencoder_tensor_info(self.handle, i, name, 64, C.byref(off), C.byref(nb), C.byref(dt)))
            lay[name.value.decode()] = (off.value, nb.value, dt.value)
        return lay

    def load_state_dict(self, sd, strict=True):
        packed = pack_zencoder(sd)
        lay = self._layout()
        if set(lay) != set(packed):
            raise _lib.ChbError("packer/library layout mismatch")
        blob = torch.zeros(self.lib.chb_zencoder_blob_bytes(self.handle), dtype=torch.uint8)
        for k, (off, nb, dt) in lay.items():
            t = packed[k].contiguous()
            want = torch.float16 if dt == _lib.F16 else torch.float32
            if t.dtype != want or t.numel() * t.element_size() != nb:
                raise _lib.ChbError("packed tensor %s has the wrong dtype/size" % k)
            blob[off:off + nb] = t.view(torch.uint8).reshape(-1)
        self.blob = blob.to(self.device)
        self.workspace = torch.empty(self.lib.chb_zencoder_workspace_bytes(self.handle) + 1024, dtype=torch.uint8,
                                     device=self.device)
        ws = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_zencoder_bind(self.handle, C.c_void_p(self.blob.data_ptr()), C.c_void_p(ws)))
        return self

    def forward(self, input, segmap):
        """Reference signature Zencoder.forward(input=img [B,3,S,S], segmap=one-hot [B,19,S,S] or labels u8 [B,S,S])."""
        if self.blob is None:
            raise _lib.ChbError("no weights loaded")
        img = input
        if not img.is_cuda:
            raise _lib.ChbError("ZencoderB200.forward takes CUDA tensors; use forward_host for host buffers")
        labels = segmap.argmax(1) if segmap.dim() == 4 else segmap
        labels = labels.to(device=self.device, dtype=torch.uint8).contiguous()
        img = img.to(torch.float32).contiguous()
        B = img.shape[0]
        if img.shape[1:] != (3, self.crop, self.crop) or labels.shape != (B, self.crop, self.crop):
            raise _lib.ChbError("bad input shapes %s %s" % (tuple(img.shape), tuple(labels.shape)))
        out = torch.empty((B, self.label_nc, 512), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_zencoder_forward(self.handle, C.c_void_p(img.data_ptr()),
                                                     C.c_void_p(labels.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                                     C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out

    def forward_host(self, img, labels):
        img = torch.as_tensor(img).to(torch.float32).contiguous()
        labels = torch.as_tensor(labels).to(torch.uint8).contiguous()
        B = img.shape[0]
        out = torch.empty((B, self.label_nc, 512), dtype=torch.float32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_zencoder_forward_host(self.handle, C.c_void_p(img.data_ptr()),
                                                          C.c_void_p(labels.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                                          C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out
