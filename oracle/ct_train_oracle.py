"""TEST INFRASTRUCTURE — CPU oracle of one colour/texture training step (config 045), torch fp32 + autograd.

Restates, with every random draw made an explicit input:
  Solver.forward             color_texture_branch/solver.py:85-117
  Solver.forward_d           solver.py:218-245  (+ forward_general_dis :186-216, WGAN-GP double backward :204-216)
  Solver.forward_g           solver.py:119-166  (+ forward_general_gen :168-184)
  train()                    my_torchlib/train_utils.py:54-89 (loss_total = sum lambda_k * loss_k, zero_grad, backward, step)
  Adam(lr 2e-4, betas .5/.999) solver.py:52-55
  loop of one step           color_texture_branch/train.py:115-148 (i = 0: D update, i = 1: G update, each on a fresh batch)
Pinned by tests/golden/ct_train_step.npz: losses and updated parameters of the unmodified reference Solver + train()
run from the same seeds (oracle/make_golden.py::main_ct_train).  Never imported by the product path.
"""
import random

import torch
import torch.nn.functional as F

from . import ct_oracle as co

# config 045 after LossUpdater.update(0) (color_texture_branch/config.py:16-39, defaults :52-96)
LAMBDAS = {"lambda_adv": 1.0, "lambda_gp": 10.0, "lambda_info": 1.0, "lambda_info_curliness": 1.0, "lambda_rec": 1000.0,
           "lambda_rgb": 0.01, "lambda_pca_std": 0.01, "lambda_moment_1": 0.01, "lambda_moment_2": 0.01,
           "lambda_cls_curliness": 0.1, "lambda_orthogonal": 0.1}
ENC_PROB = 0.3  # gan_input_from_encoder_prob
LR, BETA1, BETA2, EPS = 2e-4, 0.5, 0.999, 1e-8


def draw_randomness(B):
    """The draws of Solver.forward / forward_general_dis in the reference's own order and from the same generators
    (python `random` for the three in-place shuffles and the encoder-noise coin, torch.rand for alpha_gp)."""
    lst = list(range(B))
    random.shuffle(lst)
    p1 = list(lst)
    random.shuffle(lst)
    p2 = list(lst)
    random.shuffle(lst)
    p3 = list(lst)
    use_enc = random.random() < ENC_PROB
    return {"p1": p1, "p2": p2, "p3": p3, "use_enc": use_enc}


def curliness_predictor(sd, data):
    return co.predictor(sd, data, n_layers=3, predict=(("cls_curliness", 1),))


def forward(G, D, data, rnd):
    """solver.py:85-117.  Returns the dict of intermediate results the loss functions read."""
    d_real = co.discriminator(D, {"code": data["code"]})
    ae_mid = {"noise": d_real["noise"], "rgb_mean": data["rgb_mean"], "pca_std": data["pca_std"],
              "noise_curliness": d_real["noise_curliness"]}
    ae_out = co.eigen_generator(G, ae_mid)
    p1, p2, p3 = rnd["p1"], rnd["p2"], rnd["p3"]
    gan_in = {"rgb_mean": data["rgb_mean"][p1], "pca_std": data["pca_std"][p1],
              "noise_curliness": data["noise_curliness"][p2], "curliness_label": data["curliness_label"][p2]}
    gan_in["noise"] = d_real["noise"][p3].detach() if rnd["use_enc"] else data["noise"][p3]
    gan_mid = co.eigen_generator(G, gan_in)
    d_fake = co.discriminator(D, gan_mid)
    return {"d_real": d_real, "ae_mid": ae_mid, "ae_out": ae_out, "gan_in": gan_in, "gan_mid": gan_mid, "d_fake": d_fake}


def d_adv_direct(D, x, n_layers=4):
    for i in range(n_layers):
        x = F.leaky_relu(F.linear(x, D["net.%d.fc.weight" % i], D["net.%d.fc.bias" % i]), 0.2)
    return F.linear(x, D["net.%d.fc.weight" % n_layers], D["net.%d.fc.bias" % n_layers])[:, [0]]


def losses_d(D, data, fw, alpha):
    """solver.py:218-245 with gan_type wgan_gp (:195-196, :204-216)."""
    mse = F.mse_loss
    L = {}
    L["lambda_adv"] = torch.mean(fw["d_fake"]["adv"]) - torch.mean(fw["d_real"]["adv"])
    x_hat = (alpha * data["code"] + (1 - alpha) * fw["gan_mid"]["code"]).requires_grad_(True)
    out_hat = d_adv_direct(D, x_hat)
    dydx = torch.autograd.grad(outputs=out_hat, inputs=x_hat, grad_outputs=torch.ones_like(out_hat), retain_graph=True,
                               create_graph=True, only_inputs=True)[0]
    norm = torch.sqrt(torch.sum(dydx.reshape(dydx.shape[0], -1) ** 2, dim=1))
    L["lambda_gp"] = torch.mean((norm - 1) ** 2)
    L["lambda_info"] = mse(fw["d_fake"]["noise"], fw["gan_in"]["noise"])
    L["lambda_rec"] = mse(fw["ae_out"]["code"], data["code"])
    noise_mid = torch.cat([fw["ae_mid"]["noise_curliness"], fw["ae_mid"]["noise"]], dim=1)
    L["lambda_moment_1"] = (noise_mid.mean(dim=0) ** 2).mean()
    L["lambda_moment_2"] = (((noise_mid ** 2).mean(dim=0) - 1) ** 2).mean()
    L["lambda_info_curliness"] = mse(fw["d_fake"]["noise_curliness"], fw["gan_in"]["noise_curliness"])
    return L


def orthogonal_loss(G, n_layers=4):
    loss = 0
    for i in range(n_layers):
        U = G["subspaces.%d.U" % i]
        loss = loss + ((U @ U.t() - torch.eye(U.shape[0])) ** 2).mean()
    return loss


def losses_g(G, P, C, data, fw):
    """solver.py:119-166 (predictors frozen, eval mode; curliness_with_weight True)."""
    mse = F.mse_loss
    L = {}
    L["lambda_adv"] = -torch.mean(fw["d_fake"]["adv"])
    L["lambda_info"] = mse(fw["d_fake"]["noise"], fw["gan_in"]["noise"])
    L["lambda_rec"] = mse(fw["ae_out"]["code"], data["code"])
    p_rgb = co.predictor(P, fw["gan_mid"])
    L["lambda_rgb"] = mse(p_rgb["rgb_mean"], fw["gan_in"]["rgb_mean"])
    L["lambda_pca_std"] = mse(p_rgb["pca_std"], fw["gan_in"]["pca_std"])
    L["lambda_info_curliness"] = mse(fw["d_fake"]["noise_curliness"], fw["gan_in"]["noise_curliness"])
    cls = curliness_predictor(C, fw["gan_mid"])["cls_curliness"]
    w = fw["gan_in"]["noise_curliness"].abs()
    w = w / w.sum() * w.shape[0]
    L["lambda_cls_curliness"] = F.binary_cross_entropy(torch.sigmoid(cls), fw["gan_in"]["curliness_label"].float() / 2 + 0.5,
                                                      weight=w)
    L["lambda_orthogonal"] = orthogonal_loss(G)
    return L


def total_loss(L, lambdas=LAMBDAS):
    t = 0
    for k, v in L.items():
        t = t + v * lambdas[k]
    return t


class AdamState:
    """torch.optim.Adam semantics (no amsgrad, no weight decay) on a dict of tensors."""

    def __init__(self, params):
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, params, grads, lr=LR, b1=BETA1, b2=BETA2, eps=EPS):
        self.t += 1
        bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        for k in params:
            g = grads[k]
            self.m[k] = b1 * self.m[k] + (1 - b1) * g
            self.v[k] = b2 * self.v[k] + (1 - b2) * g * g
            denom = self.v[k].sqrt() / (bc2 ** 0.5) + eps
            params[k] = params[k] - (lr / bc1) * self.m[k] / denom


class TrainOracle:
    def __init__(self, G, D, P, C):
        self.G = {k: v.clone().float() for k, v in G.items()}
        self.D = {k: v.clone().float() for k, v in D.items()}
        self.P = {k: v.clone() for k, v in P.items()}
        self.C = {k: v.clone() for k, v in C.items()}
        self.adam_g, self.adam_d = AdamState(self.G), AdamState(self.D)

    def _leaf(self, sd):
        return {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}

    def grads_d(self, data, rnd, alpha):
        G, D = self._leaf(self.G), self._leaf(self.D)
        fw = forward(G, D, data, rnd)
        L = losses_d(D, data, fw, alpha)
        names = list(D)
        gr = torch.autograd.grad(total_loss(L), [D[k] for k in names], allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(D[k])) for k, g in zip(names, gr)}
        return {k: v.detach() for k, v in L.items()}, grads

    def grads_g(self, data, rnd):
        G, D = self._leaf(self.G), self._leaf(self.D)
        fw = forward(G, D, data, rnd)
        L = losses_g(G, self.P, self.C, data, fw)
        names = list(G)
        gr = torch.autograd.grad(total_loss(L), [G[k] for k in names], allow_unused=True)
        grads = {k: (g if g is not None else torch.zeros_like(G[k])) for k, g in zip(names, gr)}
        return {k: v.detach() for k, v in L.items()}, grads

    def step_d(self, data, rnd, alpha):
        L, g = self.grads_d(data, rnd, alpha)
        self.adam_d.step(self.D, g)
        return L

    def step_g(self, data, rnd):
        L, g = self.grads_g(data, rnd)
        self.adam_g.step(self.G, g)
        return L
