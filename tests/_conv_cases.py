"""Seeded cases for the implicit-GEMM conv operator and a plain torch fp32 reference of the same op.

Used by tests/test_conv_gpu.py (pytest -m gpu) and tools/gpu_microtest.py.
"""
import torch
import torch.nn.functional as F

from ctrlhair_b200 import ops


def _w_to_conv(w, taps, C):
    """[N, taps*C] (k = tap*C + c, tap = ky*3+kx) -> conv2d weight [N, C, kh, kw]."""
    N = w.shape[0]
    k = 3 if taps == 9 else 1
    return w.float().view(N, k, k, C).permute(0, 3, 1, 2).contiguous()


def ref_accumulate(segs):
    """fp32 conv of the fp16 operands, NCHW result [B, Nrows, H, W]."""
    acc = None
    for s in segs:
        a = s["a"]
        C = s.get("C", a.shape[3])
        off = s.get("ch_off", 0)
        taps = s.get("taps", 9)
        x = a[..., off:off + C].float().permute(0, 3, 1, 2)
        w = s["w"]
        pad = (1 if taps == 9 else 0) - s.get("a_pad", 0)  # an explicit border replaces the implicit zero pad
        if pad < 0:
            x = x[:, :, -pad:pad, -pad:pad]
            pad = 0
        if w.dim() == 3:
            y = torch.cat([F.conv2d(x[b:b + 1], _w_to_conv(w[b], taps, C), padding=pad) for b in range(x.shape[0])])
        else:
            y = F.conv2d(x, _w_to_conv(w, taps, C), padding=pad)
        acc = y if acc is None else acc + y
    return acc


def _act(v, act):
    if act == ops.ACT_RELU:
        return F.relu(v)
    if act == ops.ACT_LRELU:
        return F.leaky_relu(v, 0.2)
    if act == ops.ACT_TANH:
        return torch.tanh(v)
    return v


def ref_plain(segs, N, bias=None, bias_per_image=False, act=ops.ACT_NONE, res=None, res_shift=0):
    acc = ref_accumulate(segs)[:, :N]
    if bias is not None:
        if bias_per_image:
            acc = acc + bias[:, :N, None, None]
        else:
            acc = acc + bias[None, :N, None, None]
    if res is not None:
        r = res.permute(0, 3, 1, 2)
        if res_shift:
            r = r.repeat_interleave(2, 2).repeat_interleave(2, 3)
        acc = acc + r
    return _act(acc, act)  # NCHW fp32


def ref_modulate(segs, N, BN, bias, x, x_shift, noise, chan, act):
    acc = ref_accumulate(segs) + bias[None, :, None, None]
    half = BN // 2
    nt = N // BN
    g = torch.cat([acc[:, t * BN:t * BN + half] for t in range(nt)], 1)
    be = torch.cat([acc[:, t * BN + half:(t + 1) * BN] for t in range(nt)], 1)
    xv = x.permute(0, 3, 1, 2)
    if x_shift:
        xv = xv.repeat_interleave(2, 2).repeat_interleave(2, 3)
    a, c, nv = chan[0], chan[1], chan[2]
    xn = xv * a[None, :, None, None] + c[None, :, None, None]
    if noise is not None:
        # noise is [B, W, H]; noise[b, c, h, w] = n[b, w, h] * nv[c]  (normalization.py:111)
        xn = xn + noise.transpose(1, 2)[:, None] * nv[None, :, None, None]
    return _act(xn * (1 + g) + be, act)  # NCHW fp32


def _rand(gen, shape, scale=1.0, dtype=torch.float16, device="cuda"):
    return (torch.randn(shape, generator=gen, device="cpu") * scale).to(device=device, dtype=dtype)


def make_cases(device="cuda"):
    """Returns a list of (name, fn) where fn(impl) -> (got NCHW fp32, want NCHW fp32)."""
    cases = []

    def plain_case(name, B, H, W, seg_specs, N, BN, *, out_dtype=torch.float16, act=ops.ACT_NONE, use_bias=True,
                   use_res=False, res_shift=0, layout="nhwc", nrows=None, tile=None, seed=0, ksplit=0):
        def run(impl):
            gen = torch.Generator().manual_seed(1000 + seed)
            rows = nrows or N
            segs = []
            for (Ca, off, C, taps, per_image) in seg_specs:
                a = _rand(gen, (B, H, W, Ca), 1.0, device=device)
                K = taps * C
                wshape = (B, rows, K) if per_image else (rows, K)
                w = _rand(gen, wshape, (1.0 / K) ** 0.5, device=device)
                segs.append(dict(a=a, w=w, C=C, ch_off=off, taps=taps))
            bias = _rand(gen, (rows,), 0.5, torch.float32, device) if use_bias else None
            res = None
            if use_res:
                rh, rw = (H >> res_shift), (W >> res_shift)
                res = _rand(gen, (B, rh, rw, N), 1.0, torch.float32, device)
            got = ops.conv_igemm(segs, N, BN, bias=bias, act=act, out_dtype=out_dtype, out_layout=layout, res=res,
                                 res_shift=res_shift, tile=tile, impl=impl, ksplit=ksplit)
            got = got.float()
            if layout == "nhwc":
                got = got.permute(0, 3, 1, 2)
            want = ref_plain(segs, N, bias, False, act, res, res_shift)
            return got, want
        cases.append((name, run))

    def mod_case(name, B, H, W, C, BN, styled, *, x_shift=0, act=ops.ACT_LRELU, use_noise=True, seed=0):
        def run(impl):
            gen = torch.Generator().manual_seed(2000 + seed)
            N = 2 * C
            segs = []
            if styled:
                lab = torch.randint(0, 19, (B, H, W), generator=gen)
                oh = F.one_hot(lab, 32).to(device=device, dtype=torch.float16)
                w0 = _rand(gen, (B, N, 9 * 32), 0.1, device=device)
                segs.append(dict(a=oh, w=w0, C=32, taps=9))
            actv = F.relu(_rand(gen, (B, H, W, 256), 1.0, device=device))
            w1 = _rand(gen, (N, 9 * 128), (1.0 / (9 * 128)) ** 0.5, device=device)
            segs.append(dict(a=actv, w=w1, C=128, ch_off=128, taps=9))
            bias = _rand(gen, (N,), 0.3, torch.float32, device)
            xh, xw = H >> x_shift, W >> x_shift
            x = _rand(gen, (B, xh, xw, C), 1.0, torch.float32, device)
            noise = _rand(gen, (B, W, H), 1.0, torch.float32, device) if use_noise else None
            chan = torch.stack([torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.3,
                                torch.randn(C, generator=gen) * 0.1]).to(device)  # planar [3][C]
            got = ops.conv_igemm(segs, N, BN, bias=bias, epi=ops.EPI_MODULATE, act=act, x=x, x_shift=x_shift,
                                 noise=noise, chan=chan, impl=impl)
            got = got.float().permute(0, 3, 1, 2)
            want = ref_modulate(segs, N, BN, bias, x, x_shift, noise, chan, act)
            return got, want
        cases.append((name, run))

    # (Ca, ch_off, C, taps, per_image)
    plain_case("plain_c64_n64_16x16", 2, 16, 16, [(64, 0, 64, 9, False)], 64, 64, seed=1)
    plain_case("plain_c128_n256_res_up_lrelu_f32", 3, 32, 32, [(128, 0, 128, 9, False)], 256, 256,
               out_dtype=torch.float32, act=ops.ACT_LRELU, use_res=True, res_shift=1, seed=2)
    plain_case("plain_onehot_sw64_n128_relu", 2, 32, 32, [(32, 0, 32, 9, False)], 128, 128, act=ops.ACT_RELU, seed=3)
    plain_case("plain_conv1_plus_convs", 2, 32, 32, [(64, 0, 64, 9, False), (128, 0, 128, 1, False)], 64, 64,
               out_dtype=torch.float32, seed=4)
    plain_case("plain_r8_n512_two_ntiles", 3, 8, 8, [(128, 0, 128, 9, False)], 512, 256, out_dtype=torch.float32,
               seed=5)
    plain_case("plain_conv_img_nchw_tanh", 2, 32, 32, [(64, 0, 64, 9, False)], 3, 16, out_dtype=torch.float32,
               act=ops.ACT_TANH, layout="nchw", nrows=16, seed=6)
    plain_case("plain_persistent_many_tiles", 16, 64, 64, [(128, 0, 128, 9, False)], 256, 256, seed=7)
    plain_case("plain_deep_k_c1024", 1, 16, 16, [(1024, 0, 1024, 9, False)], 256, 256, out_dtype=torch.float32, seed=8)
    plain_case("plain_chan_window", 2, 16, 16, [(384, 128, 128, 9, False)], 128, 128, seed=9)
    plain_case("plain_ragged_24x20", 2, 24, 20, [(64, 0, 64, 9, False)], 64, 64, seed=10)
    # split-K: several CTAs per output tile (halo path, per-tap path with several images per tile, two segments with a
    # residual, a narrow N tile, more splits than CTAs can run at once is not needed: tiles * ksplit <= 148)
    plain_case("ksplit4_deep_k_c1024_halo", 1, 16, 16, [(1024, 0, 1024, 9, False)], 256, 128, out_dtype=torch.float32,
               seed=8, ksplit=4)
    plain_case("ksplit3_r8_tb2_bn32", 2, 8, 8, [(1024, 0, 1024, 9, False)], 128, 32, out_dtype=torch.float32,
               tile=(8, 8, 2), seed=21, ksplit=3)
    plain_case("ksplit2_two_segments_res_lrelu_f16", 1, 16, 16, [(256, 0, 256, 9, False), (512, 0, 512, 1, False)], 128,
               64, act=ops.ACT_LRELU, use_res=True, res_shift=1, seed=22, ksplit=2)
    plain_case("ksplit8_4x4_tb8", 8, 4, 4, [(512, 0, 512, 9, False)], 256, 64, out_dtype=torch.float32, tile=(4, 4, 8),
               seed=23, ksplit=8)
    def padded_case(name, B, H, W, C, N, BN, seed):
        def run(impl):
            gen = torch.Generator().manual_seed(3000 + seed)
            a = _rand(gen, (B, H + 2, W + 2, C), 1.0, device=device)  # e.g. a reflection-padded map
            w = _rand(gen, (N, 9 * C), (1.0 / (9 * C)) ** 0.5, device=device)
            bias = _rand(gen, (N,), 0.5, torch.float32, device)
            segs = [dict(a=a, w=w, C=C, taps=9, a_pad=1)]
            got = ops.conv_igemm(segs, N, BN, bias=bias, act=ops.ACT_TANH, out_dtype=torch.float32, impl=impl)
            return got.float().permute(0, 3, 1, 2), ref_plain(segs, N, bias, False, ops.ACT_TANH)
        cases.append((name, run))

    padded_case("plain_explicit_border_c128_n256", 2, 32, 32, 128, 256, 256, 1)
    padded_case("plain_explicit_border_c64_n64_wstat", 2, 16, 24, 64, 64, 64, 2)
    plain_case("plain_per_image_w_1x1", 5, 1, 24, [(512, 0, 512, 1, True)], 512, 256, tile=(24, 1, 1),
               act=ops.ACT_RELU, seed=11)
    plain_case("plain_tb4_rows", 6, 1, 19, [(512, 0, 512, 1, False)], 256, 256, tile=(32, 1, 4), use_bias=False,
               layout="nchw", seed=12)
    # batch-tiled 3x3 convs at tiny resolutions (shape nets: several images share one 128-row MMA tile), narrow N tiles
    plain_case("plain_tb8_3x3_r4_c512_bn64", 11, 4, 4, [(512, 0, 512, 9, False)], 256, 64, tile=(4, 4, 8),
               out_dtype=torch.float32, seed=13)
    plain_case("plain_tb32_3x3_r2_c1024_bn16", 33, 2, 2, [(1024, 0, 1024, 9, False)], 128, 16, tile=(2, 2, 32),
               out_dtype=torch.float32, seed=14)
    plain_case("plain_tb2_3x3_r8_c256_bn32", 5, 8, 8, [(256, 0, 256, 9, False)], 128, 32, tile=(8, 8, 2),
               out_dtype=torch.float32, seed=15)
    mod_case("mod_styled_c128_bn256_up", 2, 32, 32, 128, 256, True, x_shift=1, seed=1)
    mod_case("mod_unstyled_c64_bn128", 2, 32, 32, 64, 128, False, seed=2)
    mod_case("mod_styled_c256_two_ntiles_nonoise", 2, 16, 16, 256, 256, True, act=ops.ACT_NONE, use_noise=False,
             seed=3)
    return cases


def rel_err(got, want):
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-12))
