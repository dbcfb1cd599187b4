"""Host-side mirror of the reference face parser on the C ABI.

  BiSeNetB200.parsing_img      FaceParsing.parsing_img (external_code/face_parsing/my_parsing_util.py:31-47): PIL
                               bilinear resize to 512 on the host, the network + argmax on the GPU
  BiSeNetB200.get_mask         HairEditor.get_mask (hair_editor.py:331-335): label swap to the CelebAMask-HQ order
                               (my_parsing_util.py:49-54) + nearest resize to img_size, fused into the tail kernel
  pack_bisenet                 reference BiSeNet state_dict (model.py:230-254 key layout) -> blob tensors: eval BatchNorm
                               folded into every conv it follows, identity shortcuts as identity 1x1 K-segments
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

BN_EPS = 1e-5
# my_parsing_util.py:18-22 (network label order) and global_value_utils.py:49-51 (CelebAMask-HQ order used downstream)
BISENET_LABELS = ["background", "skin_other", "l_brow", "r_brow", "l_eye", "r_eye", "eye_g", "l_ear", "r_ear", "ear_r",
                  "nose", "mouth", "u_lip", "l_lip", "neck", "neck_l", "cloth", "hair", "hat"]
PARSING_LABEL_LIST = ["background", "skin_other", "nose", "eye_g", "l_eye", "r_eye", "l_brow", "r_brow", "l_ear", "r_ear",
                      "mouth", "u_lip", "l_lip", "hair", "hat", "ear_r", "neck_l", "neck", "cloth"]


def label_lut(swap=True):
    """network label i -> index of its name in PARSING_LABEL_LIST (the loop of my_parsing_util.py:49-54 as a table)."""
    if not swap:
        return np.arange(19, dtype=np.uint8)
    return np.array([PARSING_LABEL_LIST.index(n) for n in BISENET_LABELS], dtype=np.uint8)


def pil_bilinear_tables(in_size, out_size):
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c) for the BILINEAR filter: per output
    index the input window (first index, count) and its 22-bit fixed-point weights.  Python floats are C doubles and
    int() truncates like the C cast, so these are the very numbers Pillow computes."""
    import math
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coef = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - xmin
        w = []
        for x in range(n):
            a = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = 0.0
        for v in w:   # same left-to-right double sum as the C loop
            ww += v
        for x in range(n):
            v = w[x] / ww if ww != 0.0 else w[x]
            coef[xx, x] = int(-0.5 + v * (1 << 22)) if v < 0 else int(0.5 + v * (1 << 22))
        bounds[xx] = (xmin, n)
    return bounds, coef, ksize


_TABLES = {}


def resize_bilinear_u8(img, out_h, out_w):
    """PIL `Image.resize((out_w, out_h), Image.BILINEAR)` for a CUDA uint8 batch [B,H,W,C], bit exact."""
    if not (isinstance(img, torch.Tensor) and img.is_cuda and img.dtype == torch.uint8 and img.dim() == 4):
        raise _lib.ChbError("resize_bilinear_u8 takes a CUDA uint8 tensor [B,H,W,C] (there is no CPU path)")
    lib = _lib.load()
    img = img.contiguous()
    B, H, W, Cn = img.shape
    dev = img.device
    tabs = []
    for i, o in ((W, out_w), (H, out_h)):
        key = (i, o, str(dev))
        if key not in _TABLES:
            b, c, ks = pil_bilinear_tables(i, o)
            _TABLES[key] = (torch.from_numpy(b).to(dev), torch.from_numpy(c).to(dev), ks)
        tabs.append(_TABLES[key])
    tmp = torch.empty((B, H, out_w, Cn), dtype=torch.uint8, device=dev)
    out = torch.empty((B, out_h, out_w, Cn), dtype=torch.uint8, device=dev)
    (xb, xc, xks), (yb, yc, yks) = tabs
    with torch.cuda.device(dev):
        _lib.check(lib.chb_pil_resize_bilinear(
            C.c_void_p(img.data_ptr()), C.c_void_p(tmp.data_ptr()), C.c_void_p(out.data_ptr()), B, H, W, Cn, out_h, out_w,
            C.c_void_p(xb.data_ptr()), C.c_void_p(xc.data_ptr()), xks, C.c_void_p(yb.data_ptr()), C.c_void_p(yc.data_ptr()),
            yks, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def _fold(sd, conv, bn):
    """bias-free conv followed by eval BatchNorm -> (W', b'): W' = W * g/sqrt(var+eps), b' = beta - mean * g/sqrt(var+eps)."""
    w = sd[conv + ".weight"].double()
    s = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + BN_EPS)
    b = sd[bn + ".bias"].double() - sd[bn + ".running_mean"].double() * s
    return (w * s[:, None, None, None]).float(), b.float()


def _km(w, rows):
    """[N, C, kh, kw] -> fp16 [rows, kh*kw*C] (k = tap*C + c), zero rows appended up to `rows`."""
    k = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)
    out = torch.zeros((rows, k.shape[1]), dtype=torch.float32)
    out[:k.shape[0]] = k
    return out.to(torch.float16)


def _pad(v, rows):
    out = torch.zeros(rows, dtype=torch.float32)
    out[:v.shape[0]] = v
    return out


def pack_bisenet(sd, swap_labels=True):
    out = {}
    w, b = _fold(sd, "cp.resnet.conv1", "cp.resnet.bn1")          # [64, 3, 7, 7]
    sw = torch.zeros((64, 148), dtype=torch.float32)
    sw[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147)           # (ky, kx, ci)
    out["stem.w"], out["stem.b"] = sw, b
    chans = [64, 128, 256, 512]
    for li in range(4):
        Cc = chans[li]
        for bi in range(2):
            p, q = "cp.resnet.layer%d.%d" % (li + 1, bi), "layer%d.%d" % (li + 1, bi)
            w1, b1 = _fold(sd, p + ".conv1", p + ".bn1")
            w2, b2 = _fold(sd, p + ".conv2", p + ".bn2")
            out[q + ".conv1.w0"], out[q + ".conv1.b"] = _km(w1, Cc), b1
            out[q + ".conv2.w0"] = _km(w2, Cc)
            if (p + ".downsample.0.weight") in sd:                  # learned 1x1/s2 shortcut + its BatchNorm
                wd, bd = _fold(sd, p + ".downsample.0", p + ".downsample.1")
                out[q + ".conv2.w1"], out[q + ".conv2.b"] = _km(wd, Cc), b2 + bd
            else:                                                   # identity shortcut: 1.0 * x through the GEMM (exact)
                out[q + ".conv2.w1"], out[q + ".conv2.b"] = torch.eye(Cc, dtype=torch.float16), b2
    w, b = _fold(sd, "cp.conv_avg.conv", "cp.conv_avg.bn")
    out["conv_avg.w"], out["conv_avg.b"] = w.reshape(128, 512).contiguous(), b
    for arm in ("arm32", "arm16"):
        w, b = _fold(sd, "cp.%s.conv.conv" % arm, "cp.%s.conv.bn" % arm)
        out[arm + ".conv.w0"], out[arm + ".conv.b"] = _km(w, 128), b
        w, b = _fold(sd, "cp.%s.conv_atten" % arm, "cp.%s.bn_atten" % arm)
        out[arm + ".att.w"], out[arm + ".att.b"] = w.reshape(128, 128).contiguous(), b
    for head in ("conv_head32", "conv_head16"):
        w, b = _fold(sd, "cp.%s.conv" % head, "cp.%s.bn" % head)
        out[head + ".w0"], out[head + ".b"] = _km(w, 128), b
    w, b = _fold(sd, "ffm.convblk.conv", "ffm.convblk.bn")         # input = cat([feat_res8, feat_cp8]) (model.py:219)
    out["ffm.convblk.w0"], out["ffm.convblk.w1"], out["ffm.convblk.b"] = _km(w[:, :128], 256), _km(w[:, 128:], 256), b
    out["ffm.conv1.w"] = sd["ffm.conv1.weight"].float().reshape(64, 256).contiguous()
    out["ffm.conv2.w"] = sd["ffm.conv2.weight"].float().reshape(256, 64).contiguous()
    w, b = _fold(sd, "conv_out.conv.conv", "conv_out.conv.bn")
    out["conv_out.conv.w0"], out["conv_out.conv.b"] = _km(w, 256), b
    out["conv_out.conv_out.w0"] = _km(sd["conv_out.conv_out.weight"].float(), 32)
    out["conv_out.conv_out.b"] = torch.zeros(32, dtype=torch.float32)
    lut = torch.zeros(32, dtype=torch.uint8)
    lut[:19] = torch.from_numpy(label_lut(swap_labels))
    out["label_lut"] = lut.view(torch.float32)                      # 32 bytes, typed fp32 in the layout
    return out


class BiSeNetB200(torch.nn.Module):
    """Stands where `FaceParsing.bise_net` + the argmax / label-swap / resize code around it stand in the reference."""

    def __init__(self, size=512, n_classes=19, max_batch=1, device=None, swap_labels=True):
        super().__init__()
        if not torch.cuda.is_available():
            raise _lib.ChbError("BiSeNetB200 needs a CUDA device (sm_100a); there is no CPU path")
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.size, self.n_classes, self.max_batch, self.swap_labels = size, n_classes, max_batch, swap_labels
        cfg = _lib.BisenetConfig(size, n_classes, max_batch)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_check_device())
            _lib.check(self.lib.chb_bisenet_create(C.byref(cfg), C.byref(h)))
        self.handle, self.blob, self.workspace = h, None, None

    def __del__(self):
        # (at interpreter shutdown torch.nn may already be torn down: bypass nn.Module.__setattr__, never raise)
        h = self.__dict__.get("handle")
        if h:
            self.__dict__["handle"] = None
            try:
                self.lib.chb_bisenet_destroy(h)
            except Exception:
                pass

    def _layout(self):
        name = C.create_string_buffer(96)
        off, nb, dt = C.c_int64(), C.c_int64(), C.c_int()
        lay = {}
        for i in range(self.lib.chb_bisenet_num_tensors(self.handle)):
            _lib.check(self.lib.chb_bisenet_tensor_info(self.handle, i, name, 96, C.byref(off), C.byref(nb), C.byref(dt)))
            lay[name.value.decode()] = (off.value, nb.value, dt.value)
        return lay

    def load_state_dict(self, sd, strict=True):
        """Reference checkpoint format (face_parsing_79999_iter.pth: BiSeNet.state_dict())."""
        packed = pack_bisenet(sd, self.swap_labels)
        lay = self._layout()
        if set(lay) != set(packed):
            raise _lib.ChbError("packer/library layout mismatch: missing %s extra %s" %
                                (sorted(set(lay) - set(packed)), sorted(set(packed) - set(lay))))
        blob = torch.zeros(self.lib.chb_bisenet_blob_bytes(self.handle), dtype=torch.uint8)
        for k, (off, nb, dt) in lay.items():
            t = packed[k].contiguous()
            want = torch.float16 if dt == _lib.F16 else torch.float32
            if t.dtype != want or t.numel() * t.element_size() != nb:
                raise _lib.ChbError("packed tensor %s: dtype %s bytes %d, library wants %s bytes %d" %
                                    (k, t.dtype, t.numel() * t.element_size(), want, nb))
            blob[off:off + nb] = t.view(torch.uint8).reshape(-1)
        self.blob = blob.to(self.device)
        self.workspace = torch.empty(self.lib.chb_bisenet_workspace_bytes(self.handle) + 1024, dtype=torch.uint8,
                                     device=self.device)
        ws = (self.workspace.data_ptr() + 1023) // 1024 * 1024
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_bisenet_bind(self.handle, C.c_void_p(self.blob.data_ptr()), C.c_void_p(ws)))
        return self

    def forward(self, img_u8, out_size=None, return_logits=False):
        """img_u8 uint8 [B,size,size,3] RGB (CUDA) -> label map uint8 [B,out_size,out_size] (CUDA); out_size defaults
        to the network size (the full parsing map).  With return_logits also the 1/8-resolution logits [B,h,h,19]."""
        if self.blob is None:
            raise _lib.ChbError("no weights loaded")
        if not (isinstance(img_u8, torch.Tensor) and img_u8.is_cuda and img_u8.dtype == torch.uint8):
            raise _lib.ChbError("BiSeNetB200.forward takes a CUDA uint8 tensor; use forward_host for host buffers")
        img = img_u8.to(self.device).contiguous()
        B = img.shape[0]
        out_size = self.size if out_size is None else int(out_size)
        if tuple(img.shape[1:]) != (self.size, self.size, 3) or B > self.max_batch:
            raise _lib.ChbError("bad input shape %s (want [B<=%d,%d,%d,3])" % (tuple(img.shape), self.max_batch,
                                                                               self.size, self.size))
        mask = torch.empty((B, out_size, out_size), dtype=torch.uint8, device=self.device)
        h = self.size // 8
        logits = torch.empty((B, h, h, 32), dtype=torch.float32, device=self.device) if return_logits else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_bisenet_forward(
                self.handle, C.c_void_p(img.data_ptr()), C.c_void_p(mask.data_ptr()), out_size,
                C.c_void_p(logits.data_ptr()) if return_logits else None, B,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return (mask, logits[..., :self.n_classes]) if return_logits else mask

    def forward_host(self, img_u8, out_size=None):
        img = torch.as_tensor(np.ascontiguousarray(img_u8)).to(torch.uint8).contiguous()
        B = img.shape[0]
        out_size = self.size if out_size is None else int(out_size)
        if tuple(img.shape[1:]) != (self.size, self.size, 3) or B > self.max_batch:
            raise _lib.ChbError("bad input shape %s" % (tuple(img.shape),))
        mask = torch.empty((B, out_size, out_size), dtype=torch.uint8)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.chb_bisenet_forward_host(
                self.handle, C.c_void_p(img.data_ptr()), C.c_void_p(mask.data_ptr()), out_size, B,
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return mask

    # ------------------------------------------------------------------ the reference's call surface
    @staticmethod
    def resize_to_network(img_rgb, size=512):
        """my_parsing_util.py:33-35: PIL bilinear resize (third-party Pillow code; stays on the host)."""
        from PIL import Image
        return np.asarray(Image.fromarray(np.asarray(img_rgb)).resize((size, size), Image.BILINEAR))

    def parsing_img(self, img_rgb):
        """FaceParsing.parsing_img: uint8 [H,W,3] -> (parsing int64 [512,512] in the NETWORK's label order, resized image)."""
        image = self.resize_to_network(img_rgb, self.size)
        if self.swap_labels:
            raise _lib.ChbError("parsing_img returns network-order labels: build BiSeNetB200(swap_labels=False), or call "
                                "get_mask for the swapped, resized mask")
        lab = self.forward(torch.from_numpy(image[None].copy()).to(self.device))
        return lab[0].cpu().numpy().astype(np.int64), image

    def get_mask_device(self, img_rgb, img_size=256):
        """Batched HairEditor.get_mask on the device: uint8 [B,H,W,3] (host or CUDA) -> CUDA uint8 [B,img_size,img_size].
        The PIL bilinear resize to the network size runs on the GPU too (bit exact, resize_bilinear_u8)."""
        if not self.swap_labels:
            raise _lib.ChbError("get_mask needs the label swap: build BiSeNetB200(swap_labels=True)")
        img = torch.as_tensor(img_rgb).to(self.device).contiguous()
        if img.shape[1] != self.size or img.shape[2] != self.size:
            img = resize_bilinear_u8(img, self.size, self.size)
        return self.forward(img, out_size=img_size)

    def get_mask(self, img_rgb, img_size=256):
        """HairEditor.get_mask (hair_editor.py:331-335): uint8 [H,W,3] -> uint8 [img_size,img_size], CelebAMask-HQ labels.
        Also takes a batch [B,H,W,3] (-> [B,img_size,img_size]), which the reference cannot."""
        arr = np.ascontiguousarray(np.asarray(img_rgb))
        single = arr.ndim == 3
        mask = self.get_mask_device(arr[None] if single else arr, img_size).cpu().numpy()
        return mask[0] if single else mask
