"""Python front of the C-ABI operators (torch tensors in, torch tensors out; torch only provides memory + streams)."""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, EPI_MODULATE, EPI_PLAIN, F16, F32, IMPL_SIMT_DEBUG,
                   IMPL_TCGEN05, ConvDesc)


def _stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _require_cuda(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.ChbError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if dtype is not None and t.dtype != dtype:
        raise _lib.ChbError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    return t


def default_tile(H, W):
    tw = min(W, 8)
    th = min(H, 16)
    if tw * th > 128:
        th = 128 // tw
    return tw, th, 1


def conv_igemm(segs, N, BN, *, bias=None, bias_per_image=False, epi=EPI_PLAIN, act=ACT_NONE, out_dtype=torch.float16,
               out_layout="nhwc", res=None, res_shift=0, x=None, x_shift=0, noise=None, chan=None, tile=None,
               impl=IMPL_TCGEN05, out=None, ksplit=0):
    """Implicit-GEMM conv.  segs: list of dicts {a: fp16 [B,H,W,Ca], w: fp16 [Nrows,K] | [B,Nrows,K], taps, C, ch_off}.

    plain:    out = act(conv + bias (+ res[b, y>>s, x>>s, :]))            -> NHWC (or NCHW) fp16/fp32
    modulate: out = act((x*a + noise*nv + c) * (1 + gamma) + beta), fp16  -> NHWC, weight rows tiled [gamma|beta] per BN
    ksplit > 1 (plain only): that many CTAs share an output tile, each taking a slice of the channel chunks.
    """
    lib = _lib.load()
    a0 = _require_cuda(segs[0]["a"], "segs[0].a", torch.float16)
    pad0 = segs[0].get("a_pad", 0)
    B, H, W = a0.shape[0], a0.shape[1] - 2 * pad0, a0.shape[2] - 2 * pad0
    d = ConvDesc()
    d.B, d.H, d.W = B, H, W
    d.TW, d.TH, d.TB = tile if tile is not None else default_tile(H, W)
    d.nseg = len(segs)
    keep = []
    nrows = None
    for i, s in enumerate(segs):
        a = _require_cuda(s["a"], "a", torch.float16)
        w = _require_cuda(s["w"], "w", torch.float16).contiguous()
        keep += [a, w]
        if a.stride(3) != 1:
            raise _lib.ChbError("activation tensors must be channels-last contiguous in C")
        per_image = w.dim() == 3
        Cseg = s.get("C", a.shape[3])
        taps = s.get("taps", 9)
        if w.shape[-1] != taps * Cseg:
            raise _lib.ChbError("weight K (%d) != taps*C (%d)" % (w.shape[-1], taps * Cseg))
        nrows = w.shape[-2] if nrows is None else nrows
        if w.shape[-2] != nrows:
            raise _lib.ChbError("segments disagree on weight rows")
        g = d.seg[i]
        g.a = a.data_ptr()
        g.a_sb, g.a_sy, g.a_sx = a.stride(0), a.stride(1), a.stride(2)
        g.Ca, g.ch_off, g.C, g.taps = a.shape[3], s.get("ch_off", 0), Cseg, taps
        g.w = w.data_ptr()
        g.per_image = 1 if per_image else 0
        g.w_sb = 0
        g.a_pad = s.get("a_pad", 0)
    d.N, d.Nrows, d.BN = N, nrows, BN
    d.epi, d.act = epi, act
    if bias is not None:
        bias = _require_cuda(bias, "bias", torch.float32).contiguous()
        d.bias = bias.data_ptr()
        d.bias_per_image = 1 if bias_per_image else 0
    if epi == EPI_PLAIN:
        if out is None:
            shape = (B, H, W, N) if out_layout == "nhwc" else (B, N, H, W)
            out = torch.empty(shape, device=a0.device, dtype=out_dtype)
        if out_layout == "nhwc":
            d.o_sb, d.o_sy, d.o_sx, d.o_sn = out.stride(0), out.stride(1), out.stride(2), out.stride(3)
        else:
            d.o_sb, d.o_sn, d.o_sy, d.o_sx = out.stride(0), out.stride(1), out.stride(2), out.stride(3)
        d.out_dtype = F16 if out.dtype == torch.float16 else F32
        if res is not None:
            res = _require_cuda(res, "res", torch.float32)
            d.res = res.data_ptr()
            d.r_sb, d.r_sy, d.r_sx, d.r_shift = res.stride(0), res.stride(1), res.stride(2), res_shift
    else:
        x = _require_cuda(x, "x", torch.float32)
        chan = _require_cuda(chan, "chan", torch.float32).contiguous()
        Cn = N // 2
        if out is None:
            out = torch.empty((B, H, W, Cn), device=a0.device, dtype=torch.float16)
        d.o_sb, d.o_sy, d.o_sx, d.o_sn = out.stride(0), out.stride(1), out.stride(2), 1
        d.out_dtype = F16
        d.x = x.data_ptr()
        d.x_sb, d.x_sy, d.x_sx, d.x_shift = x.stride(0), x.stride(1), x.stride(2), x_shift
        d.chan = chan.data_ptr()
        if noise is not None:
            noise = _require_cuda(noise, "noise", torch.float32).contiguous()
            d.noise = noise.data_ptr()
    d.out = out.data_ptr()
    if ksplit > 1:
        ws = _ksplit_workspace(a0.device)
        d.ksplit, d.ks_ws = ksplit, ws.data_ptr()
    _lib.check(lib.chb_conv_run(C.byref(d), impl, _stream_ptr()))
    return out


_KS_WS = {}


def _ksplit_workspace(device):
    """One split-K workspace per device (counters zeroed once; every launch leaves them at zero)."""
    key = (device.type, device.index)
    if key not in _KS_WS:
        n = _lib.load().chb_conv_ksplit_workspace_bytes(1024)
        _KS_WS[key] = torch.zeros((n,), device=device, dtype=torch.uint8)
    return _KS_WS[key]


def onehot_pyramid(labels, resolutions, nclass=19):
    """labels uint8 [B,S,S] -> list of fp16 [B,r,r,32] one-hot maps (nearest-resized), bit exact."""
    lib = _lib.load()
    labels = _require_cuda(labels, "labels", torch.uint8).contiguous()
    B, S = labels.shape[0], labels.shape[1]
    outs = [torch.empty((B, r, r, 32), device=labels.device, dtype=torch.float16) for r in resolutions]
    shifts = (C.c_int * len(resolutions))(*[(S // r).bit_length() - 1 for r in resolutions])
    ptrs = (C.c_void_p * len(resolutions))(*[o.data_ptr() for o in outs])
    _lib.check(lib.chb_onehot_pyramid(labels.data_ptr(), B, S, len(resolutions), shifts, ptrs, nclass, _stream_ptr()))
    return outs


def noise_fill(n, seed, offset=0, device="cuda"):
    lib = _lib.load()
    out = torch.empty((n,), device=device, dtype=torch.float32)
    _lib.check(lib.chb_noise_fill(out.data_ptr(), n, seed, offset, _stream_ptr()))
    return out
