"""Post-processing either side of the generator on the device, with the reference's names.

  poisson_blending(source, target, mask, with_gamma)       poisson_blending.py:29-87
  postprocess_blending(face_img, res_img, face_parsing, target_parsing, blending)   hair_editor.py:257-308
  tensor_rgb_to_hsv / tensor_hsv_to_rgb                    ui/backend.py:108-125 (and :98-101)
  mask_one_hot_to_label / mask_label_to_one_hot / split_hair_face   shape_branch/shape_util.py:6-26

Inputs may be numpy arrays or tensors on any device (they are moved to the GPU); results are CUDA tensors — the
reference's callers `.cpu().numpy()` them when they need to (hair_editor.py:275).  Every function accepts a leading
batch dimension the reference does not have.  There is no CPU path.
"""
import numpy as np
import torch

from . import _lib
from .ops import _stream_ptr

HAIR_IDX = 13
DEFAULT_TOL = 1e-11      # relative residual of the conjugate-gradient solve (fp64)
DEFAULT_MAX_ITER = 6000

_LUTS = {}


def host_gamma_tables():
    """The two 256-entry tables of the gamma curve as THIS host's numpy computes them (poisson_blending.py:41-42,81):
    pow() differs in the last bit between numpy builds (SVML vs libm), which decides whether an untouched pixel of
    value v returns as v or v - 1 after `** (1/2.2) ** 2.2` and the uint8 truncation."""
    v = np.arange(256).astype("float")
    fwd = np.power(v, 1 / 2.2)
    back = np.power(fwd, 2.2)
    back[back > 255] = 255
    back[back < 0] = 0
    return fwd, back.astype("uint8")


def _device_luts(device, tables=None):
    if tables is not None:
        fwd, known = tables
        return (torch.as_tensor(np.asarray(fwd, dtype=np.float64)).to(device),
                torch.as_tensor(np.asarray(known, dtype=np.uint8)).to(device))
    key = str(device)
    if key not in _LUTS:
        fwd, known = host_gamma_tables()
        _LUTS[key] = (torch.from_numpy(fwd).to(device), torch.from_numpy(known).to(device))
    return _LUTS[key]


def _launch(device, fn, *args):
    """Every launch runs with `device` current and on ITS current stream: a BackendB200(device='cuda:1') built while
    cuda:0 is current must not launch on GPU 0 against GPU 1 pointers."""
    with torch.cuda.device(device):
        _lib.check(fn(*args, _stream_ptr(device)))


def _dev(device=None):
    if not torch.cuda.is_available():
        raise _lib.ChbError("ctrlhair_b200.blend needs a CUDA device (there is no CPU path)")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _u8(x, device):
    t = torch.as_tensor(x)
    if t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    return t.to(device).contiguous()


def image_to_u8(res_img, device=None):
    """float [B,3,H,W] (or [3,H,W]) in [-1,1] -> uint8 [B,H,W,3] (hair_editor.py:273-288)."""
    lib = _lib.load()
    device = _dev(device)
    t = torch.as_tensor(res_img).to(device=device, dtype=torch.float32)
    single = t.dim() == 3
    t = (t[None] if single else t).contiguous()
    B, _, H, W = t.shape
    out = torch.empty((B, H, W, 3), device=device, dtype=torch.uint8)
    _launch(device, lib.chb_image_to_u8, t.data_ptr(), out.data_ptr(), B, H, W)
    return out[0] if single else out


def blend_mask(target_parsing, face_parsing, device=None):
    """hair_editor.py:297-306 -> res_mask_dilated uint8 [B,H,W] (or [H,W])."""
    lib = _lib.load()
    device = _dev(device)
    tp, fp = _u8(target_parsing, device), _u8(face_parsing, device)
    single = tp.dim() == 2
    if single:
        tp, fp = tp[None], fp[None]
    tp, fp = tp.reshape(-1, tp.shape[-2], tp.shape[-1]), fp.reshape(-1, fp.shape[-2], fp.shape[-1])
    out = torch.empty_like(tp)
    _launch(device, lib.chb_blend_mask, tp.data_ptr(), fp.data_ptr(), out.data_ptr(), None, tp.shape[0], tp.shape[1],
                                  tp.shape[2])
    return out[0] if single else out


def poisson_blending(source, target, mask, with_gamma=True, tol=DEFAULT_TOL, max_iter=DEFAULT_MAX_ITER,
                     return_stats=False, device=None, gamma_tables=None):
    """source, target uint8 [H,W,3] (or [B,H,W,3]); mask [H,W] / [H,W,1] (or batched), non-zero = solve there.
    Returns uint8 [H,W,3] (or batched) on the device; with return_stats also float [B,3,2] = (iterations, residual)."""
    lib = _lib.load()
    device = _dev(device)
    s, t = _u8(source, device), _u8(target, device)
    single = s.dim() == 3
    if single:
        s, t = s[None], t[None]
    B, H, W, _ = s.shape
    m = torch.as_tensor(mask).to(device)
    m = (m != 0).to(torch.uint8).reshape(B, H, W).contiguous()
    if t.shape != s.shape or s.shape[-1] != 3:
        raise _lib.ChbError("poisson_blending: source and target must both be [.., H, W, 3]")
    out = torch.empty_like(s)
    stats = torch.zeros((B, 3, 2), device=device, dtype=torch.float32)
    fwd, known = _device_luts(device, gamma_tables)
    _launch(device, lib.chb_poisson_blend, s.data_ptr(), t.data_ptr(), m.data_ptr(), out.data_ptr(), B, H, W,
                                     1 if with_gamma else 0, float(tol), int(max_iter), stats.data_ptr(),
                                     fwd.data_ptr(), known.data_ptr())
    out = out[0] if single else out
    return (out, stats) if return_stats else out


def poisson_coarse_inverse(mask, device=None):
    """The coarse level of the solver's preconditioner (a test hook): mask [B,H,256] -> float32 [B,256,256], the inverse
    of P^T A P on 16 x 16-pixel aggregates as the solver stores it (fp16 values, un-permuted and un-scaled here)."""
    lib = _lib.load()
    device = _dev(device)
    m = torch.as_tensor(mask).to(device)
    m = (m != 0).to(torch.uint8)
    m = (m[None] if m.dim() == 2 else m).contiguous()
    B, H, W = m.shape
    if W != 256:
        raise _lib.ChbError("poisson_coarse_inverse: 256-column masks only")
    out = torch.empty((B, 256, 256), device=device, dtype=torch.float16)
    _launch(device, lib.chb_poisson_coarse_inverse, m.data_ptr(), out.data_ptr(), B, H)
    # stored column (a' & 15) * 16 + (a' >> 4) holds aggregate a'
    a = torch.arange(256, device=device)
    return out.float()[:, :, (a & 15) * 16 + (a >> 4)] / 256.0


def postprocess_blending(face_img, res_img, face_parsing, target_parsing, blending=True, tol=DEFAULT_TOL,
                         max_iter=DEFAULT_MAX_ITER, device=None, gamma_tables=None):
    """HairEditor.postprocess_blending (hair_editor.py:257-308): returns (image uint8 [H,W,3], res_mask_dilated
    [H,W,1] or None), batched when res_img is [B,3,H,W]."""
    lib = _lib.load()
    device = _dev(device)
    res = torch.as_tensor(res_img).to(device=device, dtype=torch.float32)
    single = res.dim() == 3
    res = (res[None] if single else res).contiguous()
    B, _, H, W = res.shape
    out = torch.empty((B, H, W, 3), device=device, dtype=torch.uint8)
    if not blending:
        _launch(device, lib.chb_postprocess_blending, None, res.data_ptr(), None, None, out.data_ptr(), None, None, B, H, W, 0,
                                                float(tol), int(max_iter), None, None, None)
        return (out[0] if single else out), None
    face = _u8(face_img, device).reshape(B, H, W, 3)
    fp = _u8(face_parsing, device).reshape(B, H, W)
    tp = _u8(target_parsing, device).reshape(B, H, W)
    rmd = torch.empty((B, H, W), device=device, dtype=torch.uint8)
    ws = torch.empty((int(lib.chb_postprocess_workspace_bytes(B, H, W)),), device=device, dtype=torch.uint8)
    fwd, known = _device_luts(device, gamma_tables)
    _launch(device, lib.chb_postprocess_blending, face.data_ptr(), res.data_ptr(), fp.data_ptr(), tp.data_ptr(),
                                            out.data_ptr(), rmd.data_ptr(), ws.data_ptr(), B, H, W, 1, float(tol),
                                            int(max_iter), None, fwd.data_ptr(), known.data_ptr())
    rmd = rmd[..., None]
    return (out[0], rmd[0]) if single else (out, rmd)


def tensor_rgb_to_hsv(rgb, device=None):
    """Backend.tensor_rgb_to_hsv (ui/backend.py:117-125): float or uint8 [N,3] -> uint8 HSV [N,3] (H in 0..179),
    without leaving the device.  Float input is cast like ndarray.astype('uint8') first."""
    lib = _lib.load()
    device = _dev(device)
    t = torch.as_tensor(rgb).to(device)
    shape = t.shape
    out = torch.empty(shape, device=device, dtype=torch.uint8)
    n = t.numel() // 3
    if t.dtype == torch.uint8:
        t = t.contiguous()
        _launch(device, lib.chb_rgb_to_hsv, None, t.data_ptr(), out.data_ptr(), n)
    else:
        t = t.to(torch.float32).contiguous()
        _launch(device, lib.chb_rgb_to_hsv, t.data_ptr(), None, out.data_ptr(), n)
    return out


def tensor_hsv_to_rgb(hsv, device=None):
    """Backend.tensor_hsv_to_rgb (ui/backend.py:108-115): [N,3] HSV (cast with astype('uint8') semantics) -> uint8 RGB."""
    lib = _lib.load()
    device = _dev(device)
    t = torch.as_tensor(hsv).to(device)
    if t.dtype != torch.uint8:
        t = (torch.trunc(t.to(torch.float64)).to(torch.int64) & 0xFF).to(torch.uint8)
    t = t.contiguous()
    out = torch.empty_like(t)
    _launch(device, lib.chb_hsv_to_rgb, t.data_ptr(), out.data_ptr(), t.numel() // 3)
    return out


def mask_one_hot_to_label(one_hot):
    """shape_util.py:17-20 on the device: float [B,C,H,W] -> uint8 [B,H,W] (the reference returns int64; the generator
    consumes uint8 label maps directly, so Backend.refresh_cur_mask needs no host round trip)."""
    lib = _lib.load()
    if not isinstance(one_hot, torch.Tensor) or not one_hot.is_cuda:
        raise _lib.ChbError("mask_one_hot_to_label needs a CUDA tensor (there is no CPU path)")
    t = one_hot.to(torch.float32).contiguous()
    B, Cn, H, W = t.shape
    out = torch.empty((B, H, W), device=t.device, dtype=torch.uint8)
    device = t.device
    _launch(device, lib.chb_onehot_to_label, t.data_ptr(), out.data_ptr(), B, Cn, H * W)
    return out


def mask_label_to_one_hot(img, nc=19):
    """shape_util.py:6-14 on the device: uint8 [B,1,H,W] (255 = none) -> float32 [B,19,H,W]."""
    lib = _lib.load()
    if not isinstance(img, torch.Tensor) or not img.is_cuda:
        raise _lib.ChbError("mask_label_to_one_hot needs a CUDA tensor (there is no CPU path)")
    t = img.to(torch.uint8).contiguous()
    B, _, H, W = t.shape
    out = torch.empty((B, nc, H, W), device=t.device, dtype=torch.float32)
    device = t.device
    _launch(device, lib.chb_label_to_onehot, t.data_ptr(), out.data_ptr(), B, nc, H * W)
    return out


def split_hair_face(mask):
    """shape_util.py:23-26."""
    return mask[:, [HAIR_IDX]], torch.cat([mask[:, :HAIR_IDX], mask[:, HAIR_IDX + 1:]], dim=1)
