"""Host-side weight packer: reference `netG` state_dict -> the tensors of the library's packed blob.

Load-time only (not on the hot path).  All folds are exact in eval mode:
  * spectral norm   W = weight_orig / (u^T W_mat v)                   (architecture.py:41-45, torch hook)
  * eval BatchNorm  (x - mean) * rsqrt(var + 1e-5) -> per-channel a, c (sync_batchnorm/batchnorm.py:50-55)
  * noise_var       folded with rstd                                  (normalization.py:111)
  * blending        alpha = sigmoid(blending_*) folded into conv_gamma/conv_beta (style) and (1 - alpha) into
                    mlp_gamma/mlp_beta weights, both into the bias     (normalization.py:177-182)
  * region factoring: conv_{gamma,beta}(middle_avg) == conv3x3(one_hot, Weff[b]) with
                    Weff[b][n, j, tap] = sum_ci alpha * W[n, ci, tap] * mu[b, j, ci]  (normalization.py:117-153,172-173)
Layouts are the ones csrc/generator.cu documents: conv weights [rows][tap*C + c] (tap = ky*3 + kx), one-hot
channels padded 19 -> 32, gamma/beta rows interleaved per N-tile ([gamma half | beta half]).
"""
import torch

BLOCKS = [("head_0", 16, 16, True), ("G_middle_0", 16, 16, True), ("G_middle_1", 16, 16, True),
          ("up_0", 16, 8, True), ("up_1", 8, 4, True), ("up_2", 4, 2, True), ("up_3", 2, 1, False)]
BN_EPS = 1e-5
ONEHOT_PAD = 32


def _sn(sd, p):
    w = sd[p + ".weight_orig"].float()
    sigma = torch.dot(sd[p + ".weight_u"].float(), torch.mv(w.flatten(1), sd[p + ".weight_v"].float()))
    return w / sigma


def _lo(w, wd):
    """fp16 hi+lo split: the part of fp32 `w` that rounding to fp16 drops (zeros when packing in fp32)."""
    if wd != torch.float16:
        return torch.zeros_like(w).to(wd)
    return (w.float() - w.to(torch.float16).float()).to(torch.float16)


def _k_major(w):
    """[N, C, kh, kw] -> [N, kh*kw*C] with k = tap*C + c."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _k_major_onehot(w, bias=None):
    """[N, 19, 3, 3] -> [N, 9*32], label channels zero-padded to 32.  With `bias`: the two channels after the labels
    are constant 1 in the device one-hot maps (csrc/aux_kernels.cu), and the centre tap of a zero-padded 3x3 conv
    never leaves the image, so bias[n] placed on that tap of those channels (fp16 hi + lo split: exact to 2^-22) is
    added by the GEMM itself and the epilogue has no bias to fetch."""
    n, c = w.shape[0], w.shape[1]
    out = torch.zeros((n, 3, 3, ONEHOT_PAD), dtype=w.dtype)
    out[..., :c] = w.permute(0, 2, 3, 1)
    if bias is not None:
        assert c + 2 <= ONEHOT_PAD
        hi = bias.float().to(torch.float16).float()
        out[:, 1, 1, c] = hi
        out[:, 1, 1, c + 1] = (bias.float() - hi).to(torch.float16).float()
    return out.reshape(n, 9 * ONEHOT_PAD)


def tile_gamma_beta(g, b, bn):
    """Interleaves gamma rows and beta rows per N tile of width bn: [g0..g_{h-1} | b0..b_{h-1} | g_h.. ...]."""
    C = g.shape[0]
    half = bn // 2
    nt = (2 * C) // bn
    gs = g.reshape(nt, half, *g.shape[1:])
    bs = b.reshape(nt, half, *b.shape[1:])
    return torch.cat([gs, bs], 1).reshape(2 * C, *g.shape[1:])


def ace_list(fin, fout):
    """ACE order inside a block as the schedule uses it: (ace_s,) ace_0, ace_1."""
    fmid = min(fin, fout)
    lst = [("ace_0", fin), ("ace_1", fmid)]
    if fin != fout:
        lst = [("ace_s", fin)] + lst
    return lst


def pack_generator(sd, ngf=64, label_nc=19, weight_dtype=torch.float16):
    """Returns {blob tensor name: CPU tensor} for every tensor chb_generator_tensor_info enumerates."""
    out = {}
    wd = weight_dtype
    out["fc.w"] = _k_major_onehot(sd["fc.weight"].float(), sd["fc.bias"]).to(wd)
    out["fc.b"] = sd["fc.bias"].float()
    fcmu_w, fcmu_b = [], []
    for name, fi, fo, styled in BLOCKS:
        fin, fout = fi * ngf, fo * ngf
        aces = ace_list(fin, fout)
        out[name + ".sh.w"] = torch.cat(
            [_k_major_onehot(sd["%s.%s.Spade.mlp_shared.0.weight" % (name, a)].float(),
                             sd["%s.%s.Spade.mlp_shared.0.bias" % (name, a)]) for a, _ in aces]).to(wd)
        out[name + ".sh.b"] = torch.cat([sd["%s.%s.Spade.mlp_shared.0.bias" % (name, a)].float() for a, _ in aces])
        for a, C in aces:
            p = "%s.%s" % (name, a)
            bn = min(256, 2 * C)
            wg = _k_major(sd[p + ".Spade.mlp_gamma.weight"].float())
            wb = _k_major(sd[p + ".Spade.mlp_beta.weight"].float())
            bg = sd[p + ".Spade.mlp_gamma.bias"].float()
            bb = sd[p + ".Spade.mlp_beta.bias"].float()
            if styled:
                ag = torch.sigmoid(sd[p + ".blending_gamma"].float())
                ab = torch.sigmoid(sd[p + ".blending_beta"].float())
                wg, wb = wg * (1 - ag), wb * (1 - ab)
                bg = ag * sd[p + ".conv_gamma.bias"].float() + (1 - ag) * bg
                bb = ab * sd[p + ".conv_beta.bias"].float() + (1 - ab) * bb
                # style weights: [C, 512, 3, 3] -> [C, 9, 512], rows (n, tap)
                sg = (ag * sd[p + ".conv_gamma.weight"].float()).permute(0, 2, 3, 1).reshape(C, 9, -1)
                sb = (ab * sd[p + ".conv_beta.weight"].float()).permute(0, 2, 3, 1).reshape(C, 9, -1)
                st = tile_gamma_beta(sg, sb, bn)
                out[p + ".style.w"] = st.reshape(2 * C * 9, -1).to(wd)
                fcmu_w.append(torch.stack([sd["%s.fc_mu%d.weight" % (p, j)].float() for j in range(label_nc)]))
                fcmu_b.append(torch.stack([sd["%s.fc_mu%d.bias" % (p, j)].float() for j in range(label_nc)]))
            out[p + ".gb.w"] = tile_gamma_beta(wg, wb, bn).to(wd)
            out[p + ".gb.b"] = tile_gamma_beta(bg, bb, bn)
            rstd = torch.rsqrt(sd[p + ".param_free_norm.running_var"].float() + BN_EPS)
            out[p + ".chan"] = torch.stack([rstd, -sd[p + ".param_free_norm.running_mean"].float() * rstd,
                                            sd[p + ".noise_var"].float() * rstd])  # planar [3][C]
        w0, w1 = _k_major(_sn(sd, name + ".conv_0")), _k_major(_sn(sd, name + ".conv_1"))
        out[name + ".conv_0.w"] = w0.to(wd)
        out[name + ".conv_0.b"] = sd[name + ".conv_0.bias"].float()
        out[name + ".conv_1.w"] = w1.to(wd)
        out[name + ".conv_1.b"] = sd[name + ".conv_1.bias"].float()
        out[name + ".conv_0.wlo"] = _lo(w0, wd)   # optional (CHB_PREC_W(block)): fp16 rounding residuals of the weights
        out[name + ".conv_1.wlo"] = _lo(w1, wd)
        if fin != fout:
            ws = _k_major(_sn(sd, name + ".conv_s"))
            out[name + ".conv_s.w"] = ws.to(wd)
            out[name + ".conv_s.wlo"] = _lo(ws, wd)  # optional (CHB_PREC_SHORTCUT): fp16 rounding residual of conv_s.w
    # fc_mu: [19][n_styled*512][512], class-major so that "image == class" in the grouped GEMM
    out["fcmu.w"] = torch.cat(fcmu_w, 1).to(wd)        # [19, n_styled*512, 512]
    out["fcmu.b"] = torch.cat(fcmu_b, 1)               # [19, n_styled*512]
    wi = torch.zeros((16, 9 * ngf), dtype=torch.float32)
    wi[:3] = _k_major(sd["conv_img.weight"].float())
    bi = torch.zeros(16, dtype=torch.float32)
    bi[:3] = sd["conv_img.bias"].float()
    out["conv_img.w"] = wi.to(wd)
    out["conv_img.b"] = bi
    # optional (CHB_PREC_IMG): conv_img as a 1x1 GEMM onto per-(tap, out channel) partial sums, rows tap*3 + co
    wy = torch.zeros((32, ngf), dtype=torch.float32)
    wy[:27] = sd["conv_img.weight"].float().permute(2, 3, 0, 1).reshape(27, ngf)   # [ky, kx, co, ci]
    out["conv_img.wy"] = wy.to(wd)
    out["conv_img.wylo"] = _lo(wy, wd)
    return out


OPTIONAL = (".conv_s.wlo", ".conv_0.wlo", ".conv_1.wlo", "conv_img.wy", "conv_img.wylo")


def is_optional(name):
    """Tensors only some precision policies place in the blob (chb_gen_config.precision)."""
    return name.endswith(OPTIONAL)
