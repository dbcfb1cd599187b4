"""TEST INFRASTRUCTURE — CPU oracle: a literal restatement of the reference SEAN generator forward.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; it is the checker, never the product path.

It follows the reference op for op in the *dense* form the reference executes (512-channel style map,
masked region broadcast, separate gamma/beta convolutions) — not the factored form the CUDA path uses — so
that the CUDA path is checked against the reference's algorithm and not against its own algebra:

  one-hot scatter        sean_codes/models/pix2pix_model.py:136-141
  generator forward      sean_codes/models/networks/generator.py:72-109
  SPADEResnetBlock       sean_codes/models/networks/architecture.py:69-96
  ACE                    sean_codes/models/networks/normalization.py:108-189
  SPADE                  sean_codes/models/networks/normalization.py:249-257
  eval BatchNorm         sean_codes/models/networks/sync_batchnorm/batchnorm.py:50-55
  spectral norm (eval)   torch.nn.utils.spectral_norm hook at architecture.py:41-45:  W = W_orig / (u^T W_mat v)

Parity pin: oracle/make_golden.py runs the unmodified reference modules (imported from /root/reference) on the
same seeded weights and inputs and stores their outputs under tests/golden/; tests/test_oracle.py checks this
file against those vectors.  Functions take the reference-format state_dict (keys of SPADEGenerator.state_dict()).
"""
import torch
import torch.nn.functional as F


BLOCKS = [  # name, fin/ngf, fout/ngf, styled  (generator.py:35-43; up_3 is built with use_rgb=False)
    ("head_0", 16, 16, True), ("G_middle_0", 16, 16, True), ("G_middle_1", 16, 16, True),
    ("up_0", 16, 8, True), ("up_1", 8, 4, True), ("up_2", 4, 2, True), ("up_3", 2, 1, False),
]

BN_EPS = 1e-5  # batchnorm.py:40 default eps


def one_hot(labels, nc=19):
    """pix2pix_model.py:136-141: zeros(bs, nc, h, w).scatter_(1, label, 1.0). labels: integer [B,S,S]."""
    lab = labels.long().unsqueeze(1)
    out = torch.zeros((lab.shape[0], nc, lab.shape[2], lab.shape[3]), dtype=torch.float32)
    return out.scatter_(1, lab, 1.0)


def nearest(seg, r):
    """F.interpolate(seg, size=(r, r), mode='nearest') (normalization.py:115, generator.py:75): src = floor(dst*in/out)."""
    S = seg.shape[2]
    idx = (torch.arange(r) * S) // r
    return seg[:, :, idx][:, :, :, idx]


def sn_weight(sd, prefix):
    """Eval-mode spectral norm: weight_orig / sigma, sigma = u . (W_mat v)."""
    w = sd[prefix + ".weight_orig"]
    sigma = torch.dot(sd[prefix + ".weight_u"], torch.mv(w.flatten(1), sd[prefix + ".weight_v"]))
    return w / sigma


def spade(sd, p, seg):
    """normalization.py:249-257 (its own param_free_norm is never applied)."""
    actv = F.relu(F.conv2d(seg, sd[p + ".mlp_shared.0.weight"], sd[p + ".mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"], padding=1)
    return gamma, beta


def ace(sd, p, x, seg_full, codes, noise_plane, styled):
    """normalization.py:108-189.  noise_plane: the randn(B, W, H, 1) draw of this call (or None -> zeros)."""
    B, C, H, W = x.shape
    # Part 1 (:111-112): noise, then parameter-free eval BatchNorm
    if noise_plane is not None:
        added = (noise_plane * sd[p + ".noise_var"]).transpose(1, 3)
        x = x + added
    mean = sd[p + ".param_free_norm.running_mean"][None, :, None, None]
    var = sd[p + ".param_free_norm.running_var"][None, :, None, None]
    normalized = (x - mean) / torch.sqrt(var + BN_EPS)
    # Part 2 (:115)
    seg = nearest(seg_full, H)
    if styled:
        L = sd[p + ".fc_mu0.weight"].shape[0]
        middle_avg = torch.zeros((B, L, H, W), dtype=x.dtype)
        for i in range(B):  # (:141-153)
            for j in range(seg.shape[1]):
                mask = seg[i, j].bool()
                if int(mask.sum()) > 0:
                    mu = F.relu(F.linear(codes[i, j], sd["%s.fc_mu%d.weight" % (p, j)], sd["%s.fc_mu%d.bias" % (p, j)]))
                    middle_avg[i][:, mask] = mu[:, None]
        gamma_avg = F.conv2d(middle_avg, sd[p + ".conv_gamma.weight"], sd[p + ".conv_gamma.bias"], padding=1)
        beta_avg = F.conv2d(middle_avg, sd[p + ".conv_beta.weight"], sd[p + ".conv_beta.bias"], padding=1)
        gamma_spade, beta_spade = spade(sd, p + ".Spade", seg)
        ga = torch.sigmoid(sd[p + ".blending_gamma"])
        ba = torch.sigmoid(sd[p + ".blending_beta"])
        gamma = ga * gamma_avg + (1 - ga) * gamma_spade
        beta = ba * beta_avg + (1 - ba) * beta_spade
    else:  # (:183-187)
        gamma, beta = spade(sd, p + ".Spade", seg)
    return normalized * (1 + gamma) + beta


def resblock(sd, name, x, seg, codes, noise, styled, taps=None):
    """architecture.py:69-96.  `noise` is an iterator over this forward's planes (order ace_s, ace_0, ace_1)."""
    learned = (name + ".conv_s.weight_orig") in sd
    if learned:
        x_s = ace(sd, name + ".ace_s", x, seg, codes, next(noise), styled)
        x_s = F.conv2d(x_s, sn_weight(sd, name + ".conv_s"))
    else:
        x_s = x
    h0 = ace(sd, name + ".ace_0", x, seg, codes, next(noise), styled)
    dx = F.conv2d(F.leaky_relu(h0, 0.2), sn_weight(sd, name + ".conv_0"), sd[name + ".conv_0.bias"], padding=1)
    h1 = ace(sd, name + ".ace_1", dx, seg, codes, next(noise), styled)
    dx1 = F.conv2d(F.leaky_relu(h1, 0.2), sn_weight(sd, name + ".conv_1"), sd[name + ".conv_1.bias"], padding=1)
    out = x_s + dx1
    if taps is not None:
        taps["x_" + name] = out
    return out


def up2(x):
    """nn.Upsample(scale_factor=2) (generator.py:53): nearest."""
    return x.repeat_interleave(2, 2).repeat_interleave(2, 3)


def generator_forward(sd, labels, codes, noise_planes=None, taps=None, dtype=torch.float32):
    """labels integer [B,S,S]; codes [B,19,512]; noise_planes: 18 tensors [B,W,H,1] or None. Returns [B,3,S,S]."""
    if dtype != torch.float32:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
        codes = codes.to(dtype)
        noise_planes = None if noise_planes is None else [p.to(dtype) for p in noise_planes]
    seg = one_hot(labels).to(dtype)
    S = seg.shape[2]
    sw = S // 32  # generator.py:56-70, 'normal'
    noise = iter(noise_planes if noise_planes is not None else [None] * 18)
    x = F.conv2d(nearest(seg, sw), sd["fc.weight"], sd["fc.bias"], padding=1)
    if taps is not None:
        taps["x_fc"] = x
    ups_before = {"G_middle_0", "up_0", "up_1", "up_2", "up_3"}  # generator.py:85-100
    for name, fi, fo, styled in BLOCKS:
        if name in ups_before:
            x = up2(x)
        x = resblock(sd, name, x, seg, codes, noise, styled, taps)
    x = F.conv2d(F.leaky_relu(x, 0.2), sd["conv_img.weight"], sd["conv_img.bias"], padding=1)
    return torch.tanh(x)
