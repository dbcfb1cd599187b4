"""Test helper: torch-on-CPU emulation of the *packed, region-factored* schedule of csrc/generator.cu.

It consumes ctrlhair_b200.packer output, so comparing it with the oracle checks the packer's folds, layouts and
the factoring algebra without a GPU.  With round16=True it also rounds to fp16 exactly where the kernels store
fp16 (MMA operands), which predicts the end-to-end fp16 error of the CUDA path.  Never used by the product.
"""
import torch
import torch.nn.functional as F

from ctrlhair_b200.packer import BLOCKS, ace_list


def _unpack(w, C, taps=9):
    k = 3 if taps == 9 else 1
    return w.float().reshape(w.shape[0], k, k, C).permute(0, 3, 1, 2).contiguous()


def _untile(t, bn):
    """inverse of packer.tile_gamma_beta along dim 1 (channel dim of an NCHW tensor)."""
    N = t.shape[1]
    half = bn // 2
    nt = N // bn
    v = t.reshape(t.shape[0], nt, 2, half, *t.shape[2:])
    return v[:, :, 0].reshape(t.shape[0], N // 2, *t.shape[2:]), v[:, :, 1].reshape(t.shape[0], N // 2, *t.shape[2:])


def _split16(t):
    """what an fp16 hi+lo split operand carries: hi = fp16(t), lo = fp16(t - hi)."""
    hi = t.half().float()
    return hi, (t - hi).half().float()


def emulate(packed, labels, codes, noise_planes, ngf=64, label_nc=19, round16=False, precision=0):
    """precision: chb_gen_config.precision flags (only meaningful with round16): the sites whose fp16 rounding the
    kernels compensate with an hi+lo split are modelled the same way here."""
    r16 = (lambda t: t.half().float()) if round16 else (lambda t: t)

    def rsplit(t, on):  # activation stored as [hi | lo] when `on`: both halves meet the same (fp16) weights
        if not (round16 and on):
            return r16(t)
        hi, lo = _split16(t)
        return hi + lo

    PREC_IMG, PREC_SHORTCUT = 1, 2
    B, S = labels.shape[0], labels.shape[1]
    sw = S // 32
    onehot_full = F.one_hot(labels.long(), 32).permute(0, 3, 1, 2).float()  # [B,32,S,S]

    def onehot_at(r):
        step = S // r
        return onehot_full[:, :, ::step, ::step]

    # style path
    codes16 = r16(codes.float())
    fw, fb = packed["fcmu.w"].float(), packed["fcmu.b"].float()  # [19, nS*512, 512], [19, nS*512]
    L = codes.shape[2]
    mu_all = r16(F.relu(torch.einsum("jnk,bjk->bjn", fw, codes16) + fb[None]))  # [B,19,nS*512]
    noise = iter(noise_planes if noise_planes is not None else [None] * 18)
    x = F.conv2d(onehot_at(sw), _unpack(packed["fc.w"], 32), packed["fc.b"], padding=1)
    style_idx = 0
    mults = (1, 2, 2, 4, 8, 16, 32)
    prev_r = sw
    for bidx, ((name, fi, fo, styled), mul) in enumerate(zip(BLOCKS, mults)):
        split_hs = bool(precision & PREC_SHORTCUT)
        split_h1 = bool(precision & (1 << (8 + bidx)))
        split_h0 = bool(precision & (1 << (16 + bidx)))
        split_w = bool(round16 and (precision & (1 << (24 + bidx))))   # CHB_PREC_W: one more term h_hi * w_lo
        fin, fout = fi * ngf, fo * ngf
        r = sw * mul
        if r != prev_r:
            x = x.repeat_interleave(2, 2).repeat_interleave(2, 3)
        prev_r = r
        aces = ace_list(fin, fout)
        oh = onehot_at(r)
        actv_all = r16(F.relu(F.conv2d(oh, _unpack(packed[name + ".sh.w"], 32), packed[name + ".sh.b"], padding=1)))
        hs = {}

        def modulate(ai, xin, act, split=False):
            nonlocal style_idx
            a, C = aces[ai]
            p = "%s.%s" % (name, a)
            bn = min(256, 2 * C)
            actv = actv_all[:, 128 * ai:128 * (ai + 1)]
            gb = F.conv2d(actv, _unpack(packed[p + ".gb.w"], 128), packed[p + ".gb.b"], padding=1)
            if styled:
                mu = mu_all[:, :, style_idx * L:(style_idx + 1) * L]  # [B,19,512]
                style_idx += 1
                weff = r16(torch.einsum("nk,bjk->bnj", packed[p + ".style.w"].float(), mu))  # [B, 2C*9, 19]
                weff = weff.reshape(B, 2 * C, 3, 3, label_nc).permute(0, 1, 4, 2, 3)  # [B,2C,19,3,3]
                gb = gb + torch.cat([F.conv2d(oh[b:b + 1, :label_nc], weff[b], padding=1) for b in range(B)])
            g, be = _untile(gb, bn)
            chan = packed[p + ".chan"]
            xn = xin * chan[0][None, :, None, None] + chan[1][None, :, None, None]
            nz = next(noise)
            if nz is not None:
                xn = xn + nz[..., 0].transpose(1, 2)[:, None] * chan[2][None, :, None, None]
            h = xn * (1 + g) + be
            if act:
                h = F.leaky_relu(h, 0.2)
            return h if split is None else rsplit(h, split)

        ai = 0
        if fin != fout:
            hs["s"] = modulate(0, x, False, None)
            ai = 1
        h0 = modulate(ai, x, True, split_h0)
        dx0 = F.conv2d(h0, _unpack(packed[name + ".conv_0.w"], fin), packed[name + ".conv_0.b"], padding=1)
        if split_w:
            dx0 = dx0 + F.conv2d(r16(h0), _unpack(packed[name + ".conv_0.wlo"], fin), padding=1)
        h1 = modulate(ai + 1, dx0, True, split_h1)
        out = F.conv2d(h1, _unpack(packed[name + ".conv_1.w"], min(fin, fout)), packed[name + ".conv_1.b"], padding=1)
        if split_w:
            out = out + F.conv2d(r16(h1), _unpack(packed[name + ".conv_1.wlo"], min(fin, fout)), padding=1)
        if fin != fout:
            ws = _unpack(packed[name + ".conv_s.w"], fin, 1)
            if round16 and split_hs:  # (hs_hi + hs_lo) * ws_hi + hs_hi * ws_lo
                hi, lo = _split16(hs["s"])
                out = out + F.conv2d(hi + lo, ws) + F.conv2d(hi, _unpack(packed[name + ".conv_s.wlo"], fin, 1))
            else:
                out = out + F.conv2d(r16(hs["s"]), ws)
        else:
            out = out + x
        x = out
    x = F.leaky_relu(x, 0.2)
    if round16 and (precision & PREC_IMG):
        hi, lo = _split16(x)
        wy = packed["conv_img.wy"].float()[:27].reshape(3, 3, 3, ngf).permute(2, 3, 0, 1)      # [co, ci, ky, kx]
        wl = packed["conv_img.wylo"].float()[:27].reshape(3, 3, 3, ngf).permute(2, 3, 0, 1)
        img = F.conv2d(hi + lo, wy, packed["conv_img.b"][:3], padding=1) + F.conv2d(hi, wl, padding=1)
    else:
        img = F.conv2d(r16(x), _unpack(packed["conv_img.w"], ngf)[:3], packed["conv_img.b"][:3], padding=1)
    return torch.tanh(img)
