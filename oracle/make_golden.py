"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference (from /root/reference).

Run in the build container only:   python -m oracle.make_golden
For each case it (1) checks that oracle/synth.py's key set equals SPADEGenerator.state_dict(), (2) runs the
reference modules on CPU with injected noise planes, (3) runs oracle/sean_oracle.py on the same inputs and prints
the difference, (4) stores inputs' seeds, the labels and the reference output as a golden fixture.
"""
import os
import sys
import time

import numpy as np
import torch

from . import ref_harness as rh
from . import sean_oracle as so
from ctrlhair_b200 import synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = [
    # name, crop, B, label kind, mode
    ("gen_c64_b2_blocky", 64, 2, "blocky", "batched"),
    ("gen_c64_b2_iid", 64, 2, "iid", "batched"),
    ("gen_c256_b1_blocky_ui", 256, 1, "blocky", "ui"),
]
SD_SEED, LABEL_SEED, CODE_SEED, NOISE_SEED = 1236, 1234, 1235, 1237


def main():
    torch.manual_seed(0)
    os.makedirs(OUT, exist_ok=True)
    t0 = time.time()
    sd = synth.make_state_dict(64, 19, SD_SEED)
    print("state dict: %d tensors, %.1f M params (%.1fs)" %
          (len(sd), sum(v.numel() for v in sd.values()) / 1e6, time.time() - t0))
    for name, crop, B, kind, mode in CASES:
        net = rh.build_reference_generator(sd, 64, crop)
        ref_sd = net.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), "synthetic key set differs from the reference's"
        for k in sd:
            assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
        labels = synth.make_labels(B, crop, kind, LABEL_SEED)
        codes = synth.make_codes(B, CODE_SEED)
        noise = synth.make_noise(B, crop, NOISE_SEED)
        onehot = so.one_hot(labels)
        t0 = time.time()
        if mode == "ui":
            ref = rh.run_reference_ui(net, onehot, codes[0], noise)
        else:
            ref = rh.run_reference_generator(net, onehot, codes, noise)
        t_ref = time.time() - t0
        t0 = time.time()
        taps = {}
        mine = so.generator_forward(sd, labels, codes, noise, taps=taps)
        t_or = time.time() - t0
        diff = float((mine - ref).abs().max())
        print("%s: ref %.2fs oracle %.2fs  max|oracle-ref| = %.3e  max|ref| = %.3f  std = %.3f" %
              (name, t_ref, t_or, diff, float(ref.abs().max()), float(ref.std())))
        for k, v in taps.items():
            print("   %-14s absmax %.3f std %.3f" % (k, float(v.abs().max()), float(v.std())))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), labels=labels.numpy(), out=ref.numpy(),
                            seeds=np.array([SD_SEED, LABEL_SEED, CODE_SEED, NOISE_SEED]), crop=crop,
                            oracle_diff=diff)
        del net
    print("done")


class _ADict(dict):
    """Minimal stand-in for addict.Dict (absent from the image): missing keys read as an empty, falsy dict."""

    def __getattr__(self, k):
        return self[k] if k in self else _ADict()

    def __setattr__(self, k, v):
        self[k] = v


def main_ct():
    """Golden vectors of the colour/texture MLPs from the unmodified reference modules (config 045 / p004)."""
    from . import ct_oracle as co
    g, d, pr = synth.make_ct_state_dicts()
    cfg = _ADict(lambda_rgb=0.01, lambda_pca_std=0.01, lambda_cls_curliness={0: 0.1}, curliness_dim=1,
                 g_hidden_dim=256, g_hidden_layer_num=4, SEAN_code=512, subspace_dim=2, noise_dim=8,
                 d_hidden_dim=256, d_hidden_layer_num=4, d_norm="none", d_activ="lrelu",
                 predictor=_ADict(curliness=1, rgb=1))
    pcfg = _ADict(SEAN_code=512, hidden_dim=256, hidden_layer_num=3, norm="bn", activ="lrelu", dropout=0.2,
                  predict_dict={"rgb_mean": 3, "pca_std": 1})
    with rh.reference_on_path():
        from color_texture_branch.model import Discriminator
        from color_texture_branch.model_eigengan import EigenGenerator
        from color_texture_branch.predictor.predictor_model import Predictor
        G, D, P = EigenGenerator(cfg), Discriminator(cfg), Predictor(pcfg)
    G.load_state_dict(g, strict=True)
    D.load_state_dict(d, strict=True)
    P.load_state_dict(pr, strict=True)
    G.eval(), D.eval(), P.eval()
    inp = synth.make_ct_inputs(7)
    with torch.no_grad():
        rg, rd, rp = G(inp)["code"], D({"code": inp["code"]}), P({"code": inp["code"]})
    og, od, op = co.eigen_generator(g, inp)["code"], co.discriminator(d, inp), co.predictor(pr, inp)
    print("ct: max|oracle-ref| G %.2e D %.2e P %.2e" % (
        float((rg - og).abs().max()), max(float((rd[k] - od[k]).abs().max()) for k in od),
        max(float((rp[k] - op[k]).abs().max()) for k in op)))
    np.savez_compressed(os.path.join(OUT, "ct_mlps.npz"), gen_code=rg.numpy(), dis_adv=rd["adv"].numpy(),
                        dis_noise=rd["noise"].numpy(), dis_curl=rd["noise_curliness"].numpy(),
                        pred_rgb=rp["rgb_mean"].numpy(), pred_std=rp["pca_std"].numpy(),
                        seeds=np.array([1240, 1241]), B=7)


def main_zencoder():
    """Golden style codes from the unmodified reference Zencoder (architecture.py:154-207)."""
    import warnings
    from . import zencoder_oracle as zo
    sd = synth.make_state_dict(64, 19, SD_SEED)
    net = rh.build_reference_generator(sd, 64, 256)
    for name, S, B, kind in (("zencoder_c64_b2", 64, 2, "blocky"), ("zencoder_c256_b1", 256, 1, "blocky")):
        img = synth.make_image(B, S)
        labels = synth.make_labels(B, S, kind, LABEL_SEED)
        if S == 64:
            labels[1, :, :] = 3          # image 1: two classes only -> 17 absent rows
            labels[1, :20, :] = 11
        with warnings.catch_warnings(), torch.no_grad():
            warnings.simplefilter("ignore")
            ref = net.Zencoder(input=img, segmap=so.one_hot(labels))
        mine = zo.zencoder_forward(sd, img, labels)
        print("%s: max|oracle-ref| = %.3e  max|ref| = %.3f  zero rows %d" %
              (name, float((mine - ref).abs().max()), float(ref.abs().max()), int((ref.abs().sum(2) == 0).sum())))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), labels=labels.numpy(), out=ref.numpy(), crop=S)


def main_shape():
    """Golden codes / masks from the unmodified reference shape Generator (shape_branch/model.py:146-199)."""
    from . import shape_oracle as sho
    sd = synth.make_shape_state_dict()
    cfg = _ADict(hair_dim=16, g_norm="ln", vae_hair_mode=True, pos_encoding_order=10, total_batch_size=4,
                 sample_batch_size=16)
    with rh.reference_on_path():
        from shape_branch.model import Generator
        G = Generator(cfg)
    G.load_state_dict(sd, strict=True)
    G.eval()
    hair, face = synth.make_shape_inputs(2)
    with torch.no_grad():
        hc = G.forward_hair_encoder(hair, testing=True)
        fc = G.forward_face_encoder(face)
        m = G.forward_decode_by_code(hc, fc)
    ohc, ofc = sho.forward_hair_encoder(sd, hair), sho.forward_face_encoder(sd, face)
    om = sho.forward_decode_by_code(sd, ohc, ofc)
    print("shape: max|oracle-ref| hair %.2e face %.2e mask %.2e" %
          (float((hc - ohc).abs().max()), float((fc - ofc).abs().max()), float((m - om).abs().max())))
    np.savez_compressed(os.path.join(OUT, "shape_b2.npz"), hair_code=hc.numpy(), face_code=fc.numpy(),
                        mask_argmax=m.argmax(1).to(torch.uint8).numpy(), mask_sub=m[:, :, ::8, ::8].numpy(), B=2)


def _param_sample(sd, stride=8):
    """Every `stride`-th element of every tensor, concatenated in key order (keeps the fixture small; Adam is
    elementwise, so each sampled element checks the gradient at that element independently)."""
    return np.concatenate([v.detach().reshape(-1)[::stride].numpy() for k, v in sd.items()])


def main_ct_train(B=32, n_steps=2):
    """Golden losses / updated parameters of the unmodified reference Solver + train() (config 045) over n_steps
    iterations of the loop body of color_texture_branch/train.py:115-148, from fixed seeds."""
    import random
    import types
    import warnings
    from . import ct_train_oracle as to
    addict = types.ModuleType("addict")
    addict.Dict = _ADict
    sys.modules["addict"] = addict
    g, d, pr, cu = synth.make_ct_train_state_dicts()
    argv, sys.argv = sys.argv, ["make_golden"]
    with rh.reference_on_path(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from color_texture_branch.config import cfg
        from color_texture_branch.solver import Solver
        from my_torchlib.train_utils import LossUpdater, train
        solver = Solver(cfg, torch.device("cpu"), local_rank=-1, training=True)
    sys.argv = argv
    solver.gen.load_state_dict(g, strict=True)
    solver.dis.load_state_dict(d, strict=True)
    solver.rgb_model.load_state_dict(pr, strict=True)
    solver.curliness_model.load_state_dict(cu, strict=True)
    LossUpdater(cfg).update(0)
    lam = {k: float(cfg[k]) for k in to.LAMBDAS}
    assert lam == to.LAMBDAS, lam
    assert cfg.gan_input_from_encoder_prob == to.ENC_PROB and cfg.lr_d == to.LR and cfg.beta1 == to.BETA1
    oracle = to.TrainOracle(g, d, pr, cu)
    random.seed(77)
    torch.manual_seed(78)
    st_py, st_t = random.getstate(), torch.get_rng_state()
    ref_losses, out = [], {}
    for step in range(n_steps):
        for i in range(2):
            data = synth.make_ct_train_batch(B, 1243 + 2 * step + i)
            loss_dict = {}
            solver.forward(data)
            if i == 0:
                solver.forward_d(loss_dict)
                train(cfg, loss_dict, optimizers=[solver.D_optimizer], step=step, writer=None, flag="D")
            else:
                solver.forward_g(loss_dict)
                train(cfg, loss_dict, optimizers=[solver.G_optimizer], step=step, writer=None, flag="G")
            ref_losses.append({k: float(v.detach()) for k, v in loss_dict.items()})
    # the oracle, replaying the same random streams
    random.setstate(st_py)
    torch.set_rng_state(st_t)
    worst = 0.0
    for step in range(n_steps):
        for i in range(2):
            data = synth.make_ct_train_batch(B, 1243 + 2 * step + i)
            rnd = to.draw_randomness(B)
            tag = "s%d_%s" % (step, "dg"[i])
            if i == 0:
                alpha = torch.rand(B, 1)
                L = oracle.step_d(data, rnd, alpha)
                out[tag + "_alpha"] = alpha.numpy()
            else:
                L = oracle.step_g(data, rnd)
            for k in ("p1", "p2", "p3"):
                out[tag + "_" + k] = np.array(rnd[k], dtype=np.int32)
            out[tag + "_use_enc"] = np.array(int(rnd["use_enc"]))
            rl = ref_losses[2 * step + i]
            assert set(rl) == set(L), (set(rl), set(L))
            for k in rl:
                worst = max(worst, abs(rl[k] - float(L[k])) / max(1e-6, abs(rl[k])))
            out[tag + "_loss_names"] = np.array(sorted(rl))
            out[tag + "_losses"] = np.array([rl[k] for k in sorted(rl)], dtype=np.float64)
    rg = {k: v.detach() for k, v in solver.gen.state_dict().items()}
    rd = {k: v.detach() for k, v in solver.dis.state_dict().items()}
    dg = max(float((rg[k] - oracle.G[k]).abs().max()) for k in rg)
    dd = max(float((rd[k] - oracle.D[k]).abs().max()) for k in rd)
    mv = max(float((rg[k] - g[k]).abs().max()) for k in rg), max(float((rd[k] - d[k]).abs().max()) for k in rd)
    print("ct_train: worst rel loss diff oracle-ref %.2e; max|param diff| G %.2e D %.2e (params moved by G %.2e D %.2e)"
          % (worst, dg, dd, mv[0], mv[1]))
    for s in range(2 * n_steps):
        print("   ", ref_losses[s])
    out["gen_keys"] = np.array(list(rg))
    out["dis_keys"] = np.array(list(rd))
    out["gen_sample"] = _param_sample(rg)
    out["dis_sample"] = _param_sample(rd)
    np.savez_compressed(os.path.join(OUT, "ct_train_step.npz"), B=B, n_steps=n_steps, seeds=np.array([77, 78, 1243]), **out)


if __name__ == "__main__":
    if "--shape-only" in sys.argv:
        main_shape()
    elif "--ct-only" in sys.argv:
        main_ct()
    elif "--ct-train-only" in sys.argv:
        main_ct_train()
    elif "--zencoder-only" in sys.argv:
        main_zencoder()
    else:
        main()
        main_ct()
        main_zencoder()
        main_shape()
        main_ct_train()
