"""Synthetic checkpoints and inputs in the reference's formats (used by bench.py, smoke() and the tests).

The reference ships no weights (external_model_params/ is a git-ignored download), so parity and the benchmark
use a seeded synthetic `latest_net_G.pth`-style state_dict with the exact key set / shapes of
sean_codes/models/networks/generator.py:24-53 (checked against the real module by oracle/make_golden.py).
Values are chosen to look like a trained checkpoint rather than a fresh init: converged
spectral-norm u/v, non-trivial BN running stats, noise_var != 0, blending != 0.5.
"""
import math

import torch

BLOCKS = [  # name, fin/nf, fout/nf, styled (generator.py:35-43)
    ("head_0", 16, 16, True), ("G_middle_0", 16, 16, True), ("G_middle_1", 16, 16, True),
    ("up_0", 16, 8, True), ("up_1", 8, 4, True), ("up_2", 4, 2, True), ("up_3", 2, 1, False),
]
STYLE_LEN = 512
NHIDDEN = 128


def netg_shapes(ngf=64, label_nc=19):
    """Ordered {key: shape} of SPADEGenerator.state_dict()."""
    s = {}

    def conv(name, co, ci, k, bias=True):
        s[name + ".weight"] = (co, ci, k, k)
        if bias:
            s[name + ".bias"] = (co,)

    # Zencoder (architecture.py:154-175)
    conv("Zencoder.model.1", 32, 3, 3)
    conv("Zencoder.model.4", 64, 32, 3)
    conv("Zencoder.model.7", 128, 64, 3)
    s["Zencoder.model.10.weight"] = (128, 256, 3, 3)  # ConvTranspose2d: [in, out, k, k]
    s["Zencoder.model.10.bias"] = (256,)
    conv("Zencoder.model.14", STYLE_LEN, 256, 3)
    conv("fc", 16 * ngf, label_nc, 3)
    for name, fi, fo, styled in BLOCKS:
        fin, fout = fi * ngf, fo * ngf
        fmid = min(fin, fout)

        def sn_conv(n, co, ci, k, bias):
            if bias:
                s[n + ".bias"] = (co,)
            s[n + ".weight_orig"] = (co, ci, k, k)
            s[n + ".weight_u"] = (co,)
            s[n + ".weight_v"] = (ci * k * k,)

        sn_conv(name + ".conv_0", fmid, fin, 3, True)
        sn_conv(name + ".conv_1", fout, fmid, 3, True)
        if fin != fout:
            sn_conv(name + ".conv_s", fout, fin, 1, False)
        aces = [("ace_0", fin), ("ace_1", fmid)] + ([("ace_s", fin)] if fin != fout else [])
        for an, c in aces:
            p = "%s.%s" % (name, an)
            s[p + ".blending_gamma"] = (1,)
            s[p + ".blending_beta"] = (1,)
            s[p + ".noise_var"] = (c,)
            s[p + ".Spade.param_free_norm.running_mean"] = (c,)
            s[p + ".Spade.param_free_norm.running_var"] = (c,)
            s[p + ".Spade.param_free_norm.num_batches_tracked"] = ()
            conv(p + ".Spade.mlp_shared.0", NHIDDEN, label_nc, 3)
            conv(p + ".Spade.mlp_gamma", c, NHIDDEN, 3)
            conv(p + ".Spade.mlp_beta", c, NHIDDEN, 3)
            s[p + ".param_free_norm.running_mean"] = (c,)
            s[p + ".param_free_norm.running_var"] = (c,)
            s[p + ".param_free_norm.num_batches_tracked"] = ()
            if styled:
                for j in range(label_nc):
                    s["%s.fc_mu%d.weight" % (p, j)] = (STYLE_LEN, STYLE_LEN)
                    s["%s.fc_mu%d.bias" % (p, j)] = (STYLE_LEN,)
                conv(p + ".conv_gamma", c, STYLE_LEN, 3)
                conv(p + ".conv_beta", c, STYLE_LEN, 3)
    conv("conv_img", 3, ngf, 3)
    return s


def _power_iteration(w_mat, gen, iters=40):
    """Converged u, v of torch.nn.utils.spectral_norm (left/right singular vectors of W.flatten(1))."""
    u = torch.randn(w_mat.shape[0], generator=gen)
    u = u / u.norm()
    v = None
    for _ in range(iters):
        v = torch.mv(w_mat.t(), u)
        v = v / (v.norm() + 1e-12)
        u = torch.mv(w_mat, v)
        u = u / (u.norm() + 1e-12)
    return u, v


def make_state_dict(ngf=64, label_nc=19, seed=1236):
    gen = torch.Generator().manual_seed(seed)
    shapes = netg_shapes(ngf, label_nc)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(0, dtype=torch.int64)
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shp, generator=gen) * 0.5
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shp, generator=gen) * 1.5 + 0.5
        elif k.endswith("noise_var"):
            sd[k] = torch.randn(shp, generator=gen) * 0.1
        elif k.endswith("blending_gamma") or k.endswith("blending_beta"):
            sd[k] = torch.randn(shp, generator=gen)
        elif k.endswith("weight_u") or k.endswith("weight_v"):
            sd[k] = None  # filled below
        elif k.endswith(".bias"):
            sd[k] = torch.randn(shp, generator=gen) * 0.1
        else:  # conv / linear weights
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            if "Zencoder.model.10" in k:
                fan_in = shp[0] * shp[2] * shp[3]
            if ".conv_gamma." in k or ".conv_beta." in k:
                gain = 2.0   # inputs are relu(fc_mu(.)) ~ 0.1-0.3: keep the style term comparable to SPADE's
            elif ".fc_mu" in k:
                gain = 1.5
            elif ".mlp_shared." in k:
                gain = 1.5   # one-hot input: 9 active taps of 171
            elif ".mlp_gamma." in k or ".mlp_beta." in k:
                gain = 0.7
            elif k.startswith("fc."):
                gain = 3.0
            elif k.startswith("conv_img."):
                gain = 2.0
            else:
                gain = 1.0
            sd[k] = torch.randn(shp, generator=gen) * (gain / math.sqrt(fan_in))
    for k in shapes:
        if k.endswith("weight_orig"):
            base = k[:-len("weight_orig")]
            u, v = _power_iteration(sd[k].flatten(1), gen)
            sd[base + "weight_u"] = u
            sd[base + "weight_v"] = v
    return sd


def make_labels(B, S, kind="blocky", seed=1234):
    """uint8 [B,S,S] class ids in 0..18: iid-uniform, or an 8x8 grid nearest-upsampled."""
    gen = torch.Generator().manual_seed(seed)
    if kind == "iid":
        return torch.randint(0, 19, (B, S, S), generator=gen, dtype=torch.int64).to(torch.uint8)
    g = torch.randint(0, 19, (B, 8, 8), generator=gen, dtype=torch.int64)
    rep = S // 8
    return g.repeat_interleave(rep, 1).repeat_interleave(rep, 2).to(torch.uint8)


def make_codes(B, seed=1235, zero_rows=False):
    gen = torch.Generator().manual_seed(seed)
    c = torch.randn((B, 19, STYLE_LEN), generator=gen) * 0.135
    return c


def noise_plane_shapes(B, S, ngf=64):
    """Shapes of the 18 randn(B, W, H, 1) draws of one forward, in call order (architecture.py:71-78)."""
    sw = S // 32
    out = []
    for (name, fi, fo, styled), mul in zip(BLOCKS, (1, 2, 2, 4, 8, 16, 32)):
        r = sw * mul
        n = 3 if fi != fo else 2
        out += [(B, r, r, 1)] * n
    return out


def make_noise(B, S, seed=1237):
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(shp, generator=gen) for shp in noise_plane_shapes(B, S)]


def flatten_noise(planes):
    """18 planes [B,W,H,1] -> one fp32 vector in the layout chb_generator_forward expects."""
    return torch.cat([p.reshape(-1) for p in planes])


def make_ct_state_dicts(seed=1240, code_dim=512, hidden=256, g_layers=4, d_layers=4, p_layers=3, noise_dim=8,
                        curliness_dim=1):
    """Seeded state_dicts of the colour/texture nets in the reference's key format (config 045 / predictor p004):
    (Model_G EigenGenerator, Model_D Discriminator, Predictor)."""
    gen = torch.Generator().manual_seed(seed)

    def lin(out_d, in_d, gain=1.0):
        return torch.randn((out_d, in_d), generator=gen) * (gain / math.sqrt(in_d)), torch.randn(out_d, generator=gen) * 0.1

    g = {}
    g["main_layer_in.weight"], g["main_layer_in.bias"] = lin(hidden, 3 + 1 + curliness_dim, 1.5)
    for i in range(g_layers):
        od = hidden if i < g_layers - 1 else code_dim
        g["main_layer_mid.%d.1.weight" % i], g["main_layer_mid.%d.1.bias" % i] = lin(od, hidden, 1.4)
    sub = noise_dim // g_layers
    for i in range(g_layers):
        q, _ = torch.linalg.qr(torch.randn((hidden, sub), generator=gen))
        g["subspaces.%d.U" % i] = q.t().contiguous()
        g["subspaces.%d.L" % i] = torch.tensor([3.0 * k for k in range(sub, 0, -1)]) * (1 + 0.1 * torch.randn(sub, generator=gen))
        g["subspaces.%d.mu" % i] = torch.randn(hidden, generator=gen) * 0.1
    d = {}
    out_dim = 1 + noise_dim + 1 + curliness_dim  # model.py:95-104 for config 045
    for i in range(d_layers + 1):
        in_d = code_dim if i == 0 else hidden
        od = hidden if i < d_layers else out_dim
        d["net.%d.fc.weight" % i], d["net.%d.fc.bias" % i] = lin(od, in_d, 1.4)
    pr = {}
    for i in range(p_layers + 1):
        in_d = code_dim if i == 0 else hidden
        od = hidden if i < p_layers else 4
        pr["net.%d.fc.weight" % i], pr["net.%d.fc.bias" % i] = lin(od, in_d, 1.4)
        if i < p_layers:
            pr["net.%d.norm.weight" % i] = 1 + 0.2 * torch.randn(od, generator=gen)
            pr["net.%d.norm.bias" % i] = 0.1 * torch.randn(od, generator=gen)
            pr["net.%d.norm.running_mean" % i] = 0.3 * torch.randn(od, generator=gen)
            pr["net.%d.norm.running_var" % i] = torch.rand(od, generator=gen) + 0.5
            pr["net.%d.norm.num_batches_tracked" % i] = torch.tensor(100, dtype=torch.int64)
    return g, d, pr


def make_ct_inputs(B, seed=1241):
    gen = torch.Generator().manual_seed(seed)
    return {"code": torch.randn((B, 512), generator=gen) * 0.135,
            "noise": torch.randn((B, 8), generator=gen), "noise_curliness": torch.randn((B, 1), generator=gen),
            "rgb_mean": torch.rand((B, 3), generator=gen), "pca_std": torch.rand((B, 1), generator=gen)}


def make_curliness_predictor_state_dict(seed=1242, code_dim=512, hidden=32, p_layers=3):
    """Seeded state_dict of the frozen curliness classifier (predictor config p002: hidden 32, BatchNorm1d, 1 logit)."""
    gen = torch.Generator().manual_seed(seed)
    pr = {}
    for i in range(p_layers + 1):
        in_d = code_dim if i == 0 else hidden
        od = hidden if i < p_layers else 1
        pr["net.%d.fc.weight" % i] = torch.randn((od, in_d), generator=gen) * (1.4 / math.sqrt(in_d))
        pr["net.%d.fc.bias" % i] = torch.randn(od, generator=gen) * 0.1
        if i < p_layers:
            pr["net.%d.norm.weight" % i] = 1 + 0.2 * torch.randn(od, generator=gen)
            pr["net.%d.norm.bias" % i] = 0.1 * torch.randn(od, generator=gen)
            pr["net.%d.norm.running_mean" % i] = 0.3 * torch.randn(od, generator=gen)
            pr["net.%d.norm.running_var" % i] = torch.rand(od, generator=gen) + 0.5
            pr["net.%d.norm.num_batches_tracked" % i] = torch.tensor(100, dtype=torch.int64)
    return pr


def make_ct_train_state_dicts():
    """(G, D, P_rgb, P_curliness) for the training-step tests: make_ct_state_dicts() with the subspace bases pushed
    off orthonormality, so the orthogonal regulariser (model_eigengan.py:27-31) has a non-zero gradient."""
    g, d, pr = make_ct_state_dicts()
    for k in g:
        if k.endswith(".U"):
            n = g[k].numel()
            g[k] = g[k] + 0.05 * torch.sin(torch.arange(n, dtype=torch.float32) * 0.37).reshape(g[k].shape)
    return g, d, pr, make_curliness_predictor_state_dict()


def make_ct_train_batch(B, seed=1243):
    """One training batch in the shape train.py:118-126 builds it: dataset fields (code, rgb_mean, pca_std) plus the
    per-step draws (noise, curliness_label in {-1, 1}, noise_curliness = |N(0,1)| * label)."""
    gen = torch.Generator().manual_seed(seed)
    label = (torch.randint(0, 2, (B, 1), generator=gen) * 2 - 1)
    return {"code": torch.randn((B, 512), generator=gen) * 0.135,
            "rgb_mean": torch.rand((B, 3), generator=gen), "pca_std": torch.rand((B, 1), generator=gen),
            "noise": torch.randn((B, 8), generator=gen), "curliness_label": label,
            "noise_curliness": (torch.randn((B, 1), generator=gen).abs() * label).float()}


def make_image(B, S, seed=1250):
    """fp32 [B,3,S,S] in [-1,1]: smooth low-frequency content plus noise (stands in for imgs/*.png)."""
    gen = torch.Generator().manual_seed(seed)
    low = torch.rand((B, 3, 8, 8), generator=gen) * 2 - 1
    img = torch.nn.functional.interpolate(low, size=(S, S), mode="bilinear", align_corners=False)
    return (img + 0.2 * torch.randn((B, 3, S, S), generator=gen)).clamp(-1, 1)


def shape_shapes(hair_dim=16, order=10):
    """Ordered {key: shape} of shape_branch.model.Generator.state_dict() (config 054: g_norm 'ln', VAE hair encoder)."""
    s = {}
    for name, cin, out_dim, vae in (("hair_encoder", 1, hair_dim, True), ("face_encoder", 18, 1024, False)):
        c = cin + 4 * order
        for i in range(7):
            co = min(2048, 32 * 2 ** i)
            s["%s.layers.%d.conv.weight" % (name, i)] = (co, c, 4, 4)
            s["%s.layers.%d.conv.bias" % (name, i)] = (co,)
            s["%s.layers.%d.norm.gamma" % (name, i)] = (co,)
            s["%s.layers.%d.norm.beta" % (name, i)] = (co,)
            c = co
        s[name + ".out_layer.fc.weight"] = (out_dim, 8192)
        s[name + ".out_layer.fc.bias"] = (out_dim,)
        if vae:
            s[name + ".std_out_layer.fc.weight"] = (out_dim, 8192)
            s[name + ".std_out_layer.fc.bias"] = (out_dim,)
    for name, in_dim, cout in (("hair_decoder", 1024 + hair_dim, 1), ("face_decoder", 1024, 18)):
        s[name + ".in_layer.fc.weight"] = (8192, in_dim)
        s[name + ".in_layer.fc.bias"] = (8192,)
        c = 2048
        for i in range(7):
            co = min(32 * 2 ** (6 - i), 2048)
            q = "%s.layers.%d." % (name, 2 * i + 1)
            s[q + "conv.weight"] = (co, c, 3, 3)
            s[q + "conv.bias"] = (co,)
            s[q + "norm.gamma"] = (co,)
            s[q + "norm.beta"] = (co,)
            c = co
        s[name + ".out_layer.conv.weight"] = (cout, c, 3, 3)
        s[name + ".out_layer.conv.bias"] = (cout,)
    return s


def make_shape_state_dict(seed=1260):
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in shape_shapes().items():
        if k.endswith("norm.gamma"):
            sd[k] = torch.rand(shp, generator=gen) * 0.8 + 0.4
        elif k.endswith("norm.beta"):
            sd[k] = torch.randn(shp, generator=gen) * 0.1
        elif k.endswith(".bias"):
            sd[k] = torch.randn(shp, generator=gen) * 0.1
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            sd[k] = torch.randn(shp, generator=gen) * (1.4 / math.sqrt(fan_in))
    return sd


def make_shape_inputs(B, S=256, seed=1261):
    """(hair one-hot [B,1,S,S], face one-hot [B,18,S,S]) from blocky labels, as shape_util.split_hair_face gives."""
    labels = make_labels(B, S, "blocky", seed).long()
    oh = torch.zeros((B, 19, S, S)).scatter_(1, labels[:, None], 1.0)
    return oh[:, [13]], torch.cat([oh[:, :13], oh[:, 14:]], 1)


def make_blend_case(H, W, seed):
    """Synthetic post-processing inputs (hair_editor.py:257-308): a smooth 'input face' uint8 [H,W,3], a 'generated
    image' that differs from it by a smooth offset + noise, and two label maps [H,W] (background frame touching the
    border, skin, a mouth patch, and a hair blob that moves between the input parsing and the target parsing)."""
    import numpy as np
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    base = 120 + 60 * np.sin(xx / W * 5.0)[..., None] * np.cos(yy / H * 4.0)[..., None] + g.normal(0, 12, (H, W, 3))
    face = np.clip(base, 0, 255).astype(np.uint8)
    gen = np.clip(base * 0.9 + 25 + g.normal(0, 10, (H, W, 3)), 0, 255).astype(np.uint8)

    def parsing(cx, cy, rx, ry):
        p = np.full((H, W), 1, np.uint8)
        p[(yy < H * 0.15) | (xx < W * 0.08) | (xx > W * 0.92)] = 0
        p[((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 < 1.0] = 13
        p[(np.abs(xx - W * 0.5) < W * 0.1) & (np.abs(yy - H * 0.7) < H * 0.05)] = 11
        return p
    return face, gen, parsing(W * 0.45, H * 0.25, W * 0.3, H * 0.2), parsing(W * 0.55, H * 0.3, W * 0.33, H * 0.26)


def bisenet_shapes(n_classes=19):
    """Ordered {key: shape} of external_code/face_parsing/model.py::BiSeNet(n_classes).state_dict() (checked against the
    real module by oracle/make_golden_bisenet.py): ResNet-18 context path, two attention refinement modules, feature
    fusion, three output heads."""
    s = {}

    def bn(name, c):
        s[name + ".weight"] = (c,)
        s[name + ".bias"] = (c,)
        s[name + ".running_mean"] = (c,)
        s[name + ".running_var"] = (c,)
        s[name + ".num_batches_tracked"] = ()

    def cbr(name, ci, co, k):          # ConvBNReLU (model.py:12-35)
        s[name + ".conv.weight"] = (co, ci, k, k)
        bn(name + ".bn", co)

    s["cp.resnet.conv1.weight"] = (64, 3, 7, 7)
    bn("cp.resnet.bn1", 64)
    cin = 64
    for li, co in enumerate((64, 128, 256, 512), start=1):
        for bi in range(2):
            p = "cp.resnet.layer%d.%d" % (li, bi)
            s[p + ".conv1.weight"] = (co, cin, 3, 3)
            bn(p + ".bn1", co)
            s[p + ".conv2.weight"] = (co, co, 3, 3)
            bn(p + ".bn2", co)
            if cin != co:                # stride-2 first block of layers 2-4 (resnet.py:31-36)
                s[p + ".downsample.0.weight"] = (co, cin, 1, 1)
                bn(p + ".downsample.1", co)
            cin = co
    for name, ci in (("cp.arm16", 256), ("cp.arm32", 512)):
        cbr(name + ".conv", ci, 128, 3)
        s[name + ".conv_atten.weight"] = (128, 128, 1, 1)
        bn(name + ".bn_atten", 128)
    cbr("cp.conv_head32", 128, 128, 3)
    cbr("cp.conv_head16", 128, 128, 3)
    cbr("cp.conv_avg", 512, 128, 1)
    cbr("ffm.convblk", 256, 256, 1)
    s["ffm.conv1.weight"] = (64, 256, 1, 1)
    s["ffm.conv2.weight"] = (256, 64, 1, 1)
    for name, ci, mid in (("conv_out", 256, 256), ("conv_out16", 128, 64), ("conv_out32", 128, 64)):
        cbr(name + ".conv", ci, mid, 3)
        s[name + ".conv_out.weight"] = (n_classes, mid, 1, 1)
    return s


def make_bisenet_state_dict(seed=1270, n_classes=19):
    """Seeded BiSeNet checkpoint in the reference's format (face_parsing_79999_iter.pth is a download): He-scaled conv
    weights so that activations stay O(1) through the 20-odd layers, BatchNorm with non-trivial statistics."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shp in bisenet_shapes(n_classes).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(1000, dtype=torch.int64)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(shp, generator=gen) * 1.0 + 0.5
        elif k.endswith("running_mean"):
            sd[k] = torch.randn(shp, generator=gen) * 0.2
        elif k.endswith("bn2.weight"):     # residual branch: smaller gain, the trunk does not double per block
            sd[k] = torch.rand(shp, generator=gen) * 0.3 + 0.25
        elif k.endswith("bn.weight") or k.endswith("bn1.weight") or k.endswith("bn_atten.weight") or \
                k.endswith("downsample.1.weight"):
            sd[k] = torch.rand(shp, generator=gen) * 0.6 + 0.7
        elif k.endswith(".bias"):
            sd[k] = torch.randn(shp, generator=gen) * 0.1
        else:                              # conv weights
            fan_in = shp[1] * shp[2] * shp[3]
            gain = 0.5 if k.endswith("conv_out.weight") else 1.0
            sd[k] = torch.randn(shp, generator=gen) * math.sqrt(2.0 / fan_in) * gain
    return sd
