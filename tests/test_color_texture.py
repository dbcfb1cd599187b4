"""Colour/texture MLPs: oracle vs the reference's golden vectors (CPU) and CUDA vs oracle/golden (GPU)."""
import os

import numpy as np
import pytest
import torch

from ctrlhair_b200 import synth
from oracle import ct_oracle as co

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ct_mlps.npz")
TOL = 2e-5  # fp32 on both sides; only summation order differs


def _golden():
    g = np.load(GOLD)
    return {k: torch.from_numpy(g[k]) for k in g.files if k not in ("seeds", "B")}


def test_ct_oracle_matches_reference_golden():
    g, d, pr = synth.make_ct_state_dicts()
    inp = synth.make_ct_inputs(7)
    gold = _golden()
    assert float((co.eigen_generator(g, inp)["code"] - gold["gen_code"]).abs().max()) < TOL
    od = co.discriminator(d, inp)
    assert float((od["adv"] - gold["dis_adv"]).abs().max()) < TOL
    assert float((od["noise"] - gold["dis_noise"]).abs().max()) < TOL
    assert float((od["noise_curliness"] - gold["dis_curl"]).abs().max()) < TOL
    op = co.predictor(pr, inp)
    assert float((op["rgb_mean"] - gold["pred_rgb"]).abs().max()) < TOL
    assert float((op["pca_std"] - gold["pred_std"]).abs().max()) < TOL


@pytest.mark.gpu
def test_ct_cuda_matches_oracle_and_golden():
    from ctrlhair_b200 import color_texture as ct
    g, d, pr = synth.make_ct_state_dicts()
    gold = _golden()
    G = ct.EigenGeneratorB200().load_state_dict(g)
    D = ct.CodeEncoderB200().load_state_dict(d)
    P = ct.PredictorB200().load_state_dict(pr)
    for B in (7, 1, 300):
        inp = synth.make_ct_inputs(B)
        cu = {k: v.cuda() for k, v in inp.items()}
        out_g = G(cu)["code"].cpu()
        out_d = {k: v.cpu() for k, v in D({"code": cu["code"]}).items()}
        out_p = {k: v.cpu() for k, v in P({"code": cu["code"]}).items()}
        ref_g, ref_d, ref_p = co.eigen_generator(g, inp)["code"], co.discriminator(d, inp), co.predictor(pr, inp)
        assert float((out_g - ref_g).abs().max() / ref_g.abs().max()) < TOL
        for k in ref_d:
            assert float((out_d[k] - ref_d[k]).abs().max()) < TOL * 10
        for k in ref_p:
            assert float((out_p[k] - ref_p[k]).abs().max()) < TOL * 10
        if B == 7:
            assert float((out_g - gold["gen_code"]).abs().max() / gold["gen_code"].abs().max()) < TOL
            assert float((out_p["rgb_mean"] - gold["pred_rgb"]).abs().max()) < TOL * 10
    # edit_infer (solver.py:78-83): encode, override one factor, decode
    inp = synth.make_ct_inputs(4)
    cu = {k: v.cuda() for k, v in inp.items()}
    code = ct.edit_infer(D, G, cu["code"], {"rgb_mean": cu["rgb_mean"], "pca_std": cu["pca_std"]}).cpu()
    inner = co.discriminator(d, inp)
    inner.update(rgb_mean=inp["rgb_mean"], pca_std=inp["pca_std"])
    ref = co.eigen_generator(g, inner)["code"]
    assert float((code - ref).abs().max() / ref.abs().max()) < TOL * 5


@pytest.mark.gpu
def test_ct_strict_state_dict_and_host_tensors_rejected():
    from ctrlhair_b200 import _lib
    from ctrlhair_b200 import color_texture as ct
    g, d, pr = synth.make_ct_state_dicts()
    bad = dict(g)
    bad.pop("subspaces.0.mu")
    with pytest.raises(RuntimeError):
        ct.EigenGeneratorB200().load_state_dict(bad)
    P = ct.PredictorB200().load_state_dict(pr)
    with pytest.raises(_lib.ChbError):
        P({"code": torch.zeros(2, 512)})
