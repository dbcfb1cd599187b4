"""Colour/texture training step: oracle vs the reference Solver's golden step (CPU), CUDA step vs the
oracle's autograd gradients and vs the golden losses / updated parameters (GPU), host logic of the solver mirror."""
import os
import random

import numpy as np
import pytest
import torch

from ctrlhair_b200 import synth
from oracle import ct_oracle as co
from oracle import ct_train_oracle as to

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ct_train_step.npz")
LOSS_RTOL = 2e-4   # fp32 both sides, different summation order; lambda_orthogonal is a sum of near-cancelling terms
GRAD_RTOL = 2e-4   # max|dg| / max|g| per tensor
PARAM_ATOL = 2e-6  # parameters move by ~2e-4 per Adam step


def _gold():
    return np.load(GOLD)


def _rnd(g, tag):
    return {"p1": g[tag + "_p1"].tolist(), "p2": g[tag + "_p2"].tolist(), "p3": g[tag + "_p3"].tolist(),
            "use_enc": bool(int(g[tag + "_use_enc"]))}


def _sample(sd, keys, stride=8):
    return np.concatenate([sd[str(k)].detach().reshape(-1)[::stride].cpu().numpy() for k in keys])


def test_train_oracle_matches_reference_golden():
    g = _gold()
    B, n_steps = int(g["B"]), int(g["n_steps"])
    G, D, P, Cp = synth.make_ct_train_state_dicts()
    orc = to.TrainOracle(G, D, P, Cp)
    for step in range(n_steps):
        for i in range(2):
            tag = "s%d_%s" % (step, "dg"[i])
            data = synth.make_ct_train_batch(B, 1243 + 2 * step + i)
            L = orc.step_d(data, _rnd(g, tag), torch.from_numpy(g[tag + "_alpha"])) if i == 0 else \
                orc.step_g(data, _rnd(g, tag))
            names, vals = [str(n) for n in g[tag + "_loss_names"]], g[tag + "_losses"]
            assert sorted(L) == names
            for n, v in zip(names, vals):
                assert abs(float(L[n]) - v) <= 1e-5 * max(1.0, abs(v)), (tag, n, float(L[n]), v)
    assert np.abs(_sample(orc.G, g["gen_keys"]) - g["gen_sample"]).max() < 1e-7
    assert np.abs(_sample(orc.D, g["dis_keys"]) - g["dis_sample"]).max() < 1e-7


def test_randomness_is_drawn_like_the_reference():
    """The product's draw order (solver.py:98-111) equals the oracle's, which the golden run pinned to the reference."""
    from ctrlhair_b200 import ct_train
    fake = ct_train.SolverB200.__new__(ct_train.SolverB200)
    fake.B, fake.enc_prob = 32, 0.3
    random.seed(77)
    mine = [ct_train.SolverB200.draw_randomness(fake) for _ in range(4)]
    random.seed(77)
    ref = [to.draw_randomness(32) for _ in range(4)]
    assert mine == ref
    g = _gold()
    assert mine[0]["p1"] == g["s0_d_p1"].tolist() and mine[1]["p3"] == g["s0_g_p3"].tolist()
    fake.handle = None


def test_fold_predictor_equals_eval_batchnorm():
    from ctrlhair_b200 import ct_train
    _, _, P, Cp = synth.make_ct_train_state_dicts()
    x = {"code": synth.make_ct_train_batch(9)["code"]}
    for sd, predict in ((P, (("rgb_mean", 3), ("pca_std", 1))), (Cp, (("cls", 1),))):
        want = torch.cat(list(co.predictor(sd, x, predict=predict).values()), dim=1)
        f = ct_train.fold_predictor(sd)
        h = x["code"]
        for i in range(3):
            h = torch.nn.functional.leaky_relu(torch.nn.functional.linear(h, f["net.%d.fc.weight" % i], f["net.%d.fc.bias" % i]), 0.2)
        got = torch.nn.functional.linear(h, f["net.3.fc.weight"], f["net.3.fc.bias"])
        assert float((got - want).abs().max()) < 2e-6


def _make_solver(B, use_graph=True):
    from ctrlhair_b200 import ct_train
    G, D, P, Cp = synth.make_ct_train_state_dicts()
    s = ct_train.SolverB200(None, "cuda", batch_size=B, use_graph=use_graph)
    s.load_state_dicts(G, D, P, Cp)
    return s, (G, D, P, Cp)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _kink_margin(fn):
    """Smallest |pre-activation| any leaky-ReLU sees while fn() runs.  The losses (the gradient penalty above all) and
    the gradients are discontinuous where a pre-activation crosses zero, so two correct fp32 implementations can
    legitimately land on different sides when one sits within rounding distance of it."""
    import torch.nn.functional as F
    orig, seen = F.leaky_relu, []

    def spy(x, *a, **k):
        seen.append(float(x.detach().abs().min()))
        return orig(x, *a, **k)
    F.leaky_relu = spy
    try:
        out = fn()
    finally:
        F.leaky_relu = orig
    return min(seen), out


def _well_conditioned_batch(orc, B, use_enc, alpha, first_seed=4321, margin=4e-6):
    """First seeded batch whose leaky-ReLU pre-activations all stay `margin` away from zero in the oracle (fp32 rounding
    of these pre-activations is ~3e-7), so that the comparison below tests arithmetic and not which side of a kink a rounding error fell."""
    for seed in range(first_seed, first_seed + 64):
        data = synth.make_ct_train_batch(B, seed)
        random.seed(5)
        rnd = to.draw_randomness(B)
        rnd["use_enc"] = use_enc
        m_d, d_res = _kink_margin(lambda: orc.grads_d(data, rnd, alpha))
        m_g, g_res = _kink_margin(lambda: orc.grads_g(data, rnd))
        if min(m_d, m_g) > margin:
            return data, rnd, d_res, g_res
    raise AssertionError("no well-conditioned batch found")


@pytest.mark.gpu
@pytest.mark.parametrize("B,use_enc", [(32, False), (32, True), (50, False)])
def test_cuda_gradients_match_oracle_autograd(B, use_enc):
    s, (G, D, P, Cp) = _make_solver(B)
    orc = to.TrainOracle(G, D, P, Cp)
    alpha = torch.rand(B, 1, generator=torch.Generator().manual_seed(9))
    data, rnd, (Ld, gd), (Lg, gg) = _well_conditioned_batch(orc, B, use_enc, alpha)
    # discriminator sub-step
    s.forward(data, rnd)
    ld = {}
    s.forward_d(ld, alpha)
    s.synchronize()
    for k, v in Ld.items():
        assert abs(float(ld[k]) - float(v)) <= LOSS_RTOL * max(1e-3, abs(float(v))), ("D", k, float(ld[k]), float(v))
    mine = s.dis_grads(like=D)
    for k in gd:
        assert _rel(mine[k].cpu(), gd[k]) < GRAD_RTOL or float(gd[k].abs().max()) == 0.0, ("D", k, _rel(mine[k].cpu(), gd[k]))
    # generator sub-step (same parameters: no optimizer step was taken)
    lg = {}
    s.forward_g(lg)
    s.synchronize()
    for k, v in Lg.items():
        assert abs(float(lg[k]) - float(v)) <= LOSS_RTOL * max(1e-3, abs(float(v))), ("G", k, float(lg[k]), float(v))
    mine = s.gen_grads(like=G)
    for k in gg:
        assert _rel(mine[k].cpu(), gg[k]) < GRAD_RTOL, ("G", k, _rel(mine[k].cpu(), gg[k]))
    # the default form: an explicit graph of > 50 operation nodes per sub-step with more dependency edges than nodes
    assert s.launches(0) > 50 and s.launches(1) > 50
    assert s.schedule(0)[1] > s.schedule(0)[0] and s.schedule(1)[1] > s.schedule(1)[0]


@pytest.mark.gpu
def test_cuda_training_steps_match_reference_golden():
    """Two iterations of the train.py loop body (D, G, D, G) with the reference's seeds: losses and updated parameters
    against what the unmodified reference Solver + train() produced."""
    from ctrlhair_b200 import ct_train
    g = _gold()
    B, n_steps = int(g["B"]), int(g["n_steps"])
    s, (G, D, _, _) = _make_solver(B)
    random.seed(77)
    torch.manual_seed(78)
    for step in range(n_steps):
        for i in range(2):
            tag = "s%d_%s" % (step, "dg"[i])
            data = synth.make_ct_train_batch(B, 1243 + 2 * step + i)
            loss_dict = {}
            s.forward(data)                       # draws the shuffles / coin from python `random` like solver.py
            if i == 0:
                s.forward_d(loss_dict)            # draws alpha_gp from torch.rand like solver.py:199
                ct_train.train(s.cfg, loss_dict, optimizers=[s.D_optimizer])
            else:
                s.forward_g(loss_dict)
                ct_train.train(s.cfg, loss_dict, optimizers=[s.G_optimizer])
            s.synchronize()
            assert s._rnd["p1"] == g[tag + "_p1"].tolist() and s._rnd["use_enc"] == bool(int(g[tag + "_use_enc"]))
            for n, v in zip([str(n) for n in g[tag + "_loss_names"]], g[tag + "_losses"]):
                assert abs(float(loss_dict[n]) - v) <= LOSS_RTOL * max(1e-3, abs(v)), (tag, n, float(loss_dict[n]), v)
    for sd, keys, want, like in ((s.gen_state_dict(G), g["gen_keys"], g["gen_sample"], G),
                                 (s.dis_state_dict(D), g["dis_keys"], g["dis_sample"], D)):
        got = _sample(sd, keys)
        moved = np.abs(_sample(like, keys) - want)
        diff = np.abs(got - want)
        assert moved.max() > 3e-4                      # the fixture really contains two Adam steps
        assert diff.max() < 4.5e-4 and np.mean(diff > PARAM_ATOL) < 2e-3, (diff.max(), np.mean(diff > PARAM_ATOL))


@pytest.mark.gpu
def test_persistent_kernel_equals_graph_and_plain_launches():
    """The three ways a sub-step reaches the GPU run the same operations.  The dependency graph (independent operations
    side by side, every accumulation into a shared buffer kept in program order by its write-after-write edge) and plain
    launches are bitwise equal; the persistent kernel runs the two loss reductions with 256 instead of 1024 threads (a
    different but equally valid fp32 summation order), so it is held to round-off."""
    B = 32
    outs = []
    for use_graph in ("graph", False, "persistent"):
        s, _ = _make_solver(B, use_graph=use_graph)
        random.seed(3)
        torch.manual_seed(4)
        for it in range(3):   # iteration 0 captures, 1-2 replay
            data = synth.make_ct_train_batch(B, 100 + it)
            ld = {}
            s.forward(data)
            s.forward_d(ld)
            s.D_optimizer.step()
            s.forward(data)
            lg = {}
            s.forward_g(lg)
            s.G_optimizer.step()
        s.synchronize()
        outs.append((s.state.clone(), float(ld["total"]), float(lg["total"])))
        if use_graph == "persistent":
            (nd, bd), (ng, bg) = s.schedule(0), s.schedule(1)
            assert 0 < bd < nd and 0 < bg < ng        # independent operations share a phase: fewer barriers than operations
            assert s.launches(0) == 2 and s.launches(1) == 2
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1:] == outs[1][1:]
    # after three iterations (six Adam updates) the parameters agree to round-off
    n_par = outs[0][0].numel()
    d = (outs[2][0] - outs[0][0]).abs()
    assert float(d.max()) < 2e-5, float(d.max())
    assert abs(outs[2][1] - outs[0][1]) <= 1e-5 * abs(outs[0][1]) and abs(outs[2][2] - outs[0][2]) <= 1e-5 * abs(outs[0][2])
    assert n_par > 0


@pytest.mark.gpu
def test_solver_has_no_cpu_path():
    from ctrlhair_b200 import _lib, ct_train
    with pytest.raises(_lib.ChbError):
        ct_train.SolverB200(None, "cpu", batch_size=8)
