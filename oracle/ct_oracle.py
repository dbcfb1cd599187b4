"""TEST INFRASTRUCTURE — CPU oracle of the colour/texture MLPs (restated from the reference, torch fp32).

  EigenGenerator   color_texture_branch/model_eigengan.py:14-31 (SubspaceLayer), :62-83 (forward)
  Discriminator    color_texture_branch/model.py:86-127 (config 045: norm none, lrelu 0.2)
  Predictor        color_texture_branch/predictor/predictor_model.py:14-41, LinearBlock my_torchlib/module.py:56-64,
                   eval-mode BatchNorm1d, dropout inactive
Pinned by tests/golden/ct_mlps.npz (outputs of the unmodified reference modules, oracle/make_golden.py).
"""
import torch
import torch.nn.functional as F


def eigen_generator(sd, data, n_layers=4, subspace_dim=2):
    noise = data["noise"].reshape(len(data["noise"]), n_layers, subspace_dim)
    x = torch.cat([data["noise_curliness"], data["rgb_mean"], data["pca_std"]], dim=1)
    x = F.linear(x, sd["main_layer_in.weight"], sd["main_layer_in.bias"])
    for i in range(n_layers):
        sub = (sd["subspaces.%d.L" % i] * noise[:, i, :]) @ sd["subspaces.%d.U" % i] + sd["subspaces.%d.mu" % i]
        x = x + sub
        x = F.linear(F.leaky_relu(x, 0.2), sd["main_layer_mid.%d.1.weight" % i], sd["main_layer_mid.%d.1.bias" % i])
    return {"code": x}


def discriminator(sd, data, n_layers=4, noise_dim=8, curliness_dim=1):
    x = data["code"]
    for i in range(n_layers):
        x = F.leaky_relu(F.linear(x, sd["net.%d.fc.weight" % i], sd["net.%d.fc.bias" % i]), 0.2)
    out = F.linear(x, sd["net.%d.fc.weight" % n_layers], sd["net.%d.fc.bias" % n_layers])
    p = 1 + noise_dim
    return {"adv": out[:, [0]], "noise": out[:, 1:p], "noise_curliness": out[:, p:p + curliness_dim]}


def predictor(sd, data, n_layers=3, predict=(("rgb_mean", 3), ("pca_std", 1))):
    x = data["code"]
    for i in range(n_layers):
        x = F.linear(x, sd["net.%d.fc.weight" % i], sd["net.%d.fc.bias" % i])
        x = F.batch_norm(x, sd["net.%d.norm.running_mean" % i], sd["net.%d.norm.running_var" % i],
                         sd["net.%d.norm.weight" % i], sd["net.%d.norm.bias" % i], False, 0.1, 1e-5)
        x = F.leaky_relu(x, 0.2)
    out = F.linear(x, sd["net.%d.fc.weight" % n_layers], sd["net.%d.fc.bias" % n_layers])
    res, p = {}, 0
    for k, d in predict:
        res[k] = out[:, p:p + d]
        p += d
    return res
