"""Parity of the CUDA generator (C ABI -> tcgen05 kernels) against the oracle and the reference's golden vectors.

Tolerances (written out, as the task asks):
  * integer work (one-hot scatter / nearest pyramid): bit exact (tests/test_aux_gpu.py)
  * generator image, fp16 tensor-core operands with fp32 accumulation, vs the fp32 reference:
      rel-L2  ||d|| / ||ref||            <= 1e-3   (north_star's "1e-3 relative")
      max-norm max|d| / max|ref|         <= 1e-3   (met by the default "parity" precision policy —
                                                    fp16 hi+lo split of conv_img, conv_s and conv_1 inputs; the
                                                    single-pass "fast" policy sits at 1.5-1.8e-3: DESIGN.md numerics)
"""
import os

import numpy as np
import pytest
import torch

from ctrlhair_b200 import _lib
from ctrlhair_b200.generator import SeanGeneratorB200
from oracle import sean_oracle as so
from ctrlhair_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
L2_TOL = 1e-3
MAX_TOL = 1e-3


def _errs(got, ref):
    d = got - ref
    return float(d.norm() / ref.norm()), float(d.abs().max() / ref.abs().max())


@pytest.fixture(scope="module")
def gen64(synthetic_sd):
    g = SeanGeneratorB200(crop=64, max_batch=4)
    g.load_state_dict(synthetic_sd)
    return g


@pytest.fixture(scope="module")
def gen256(synthetic_sd):
    g = SeanGeneratorB200(crop=256, max_batch=8)
    g.load_state_dict(synthetic_sd)
    return g


@pytest.mark.parametrize("kind", ["blocky", "iid"])
def test_generator_vs_oracle_and_golden_c64(synthetic_sd, gen64, kind):
    B = 2
    labels, codes, noise = synth.make_labels(B, 64, kind), synth.make_codes(B), synth.make_noise(B, 64)
    ref = so.generator_forward(synthetic_sd, labels, codes, noise)
    gold = torch.from_numpy(np.load(os.path.join(GOLD, "gen_c64_b2_%s.npz" % kind))["out"])
    out = gen64.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(noise).cuda()).cpu()
    assert torch.isfinite(out).all()
    for want in (ref, gold):
        l2, mx = _errs(out, want)
        assert l2 < L2_TOL and mx < MAX_TOL, (l2, mx)


@pytest.mark.parametrize("B", [4, 5])
def test_generator_small_batch_policy_boundary_vs_oracle(synthetic_sd, B):
    """B = 4 still takes the interactive schedule (extra h_0 split, split-K where it pays), B = 5 the batch schedule:
    both against the oracle, every image."""
    g = SeanGeneratorB200(crop=64, max_batch=B)
    g.load_state_dict(synthetic_sd)
    labels, codes, noise = synth.make_labels(B, 64, "iid", seed=40 + B), synth.make_codes(B, seed=50 + B), synth.make_noise(B, 64)
    ref = so.generator_forward(synthetic_sd, labels, codes, noise)
    out = g.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(noise).cuda()).cpu()
    for i in range(B):
        l2, mx = _errs(out[i:i + 1], ref[i:i + 1])
        assert l2 < L2_TOL and mx < MAX_TOL, (i, l2, mx)


def test_generator_vs_reference_golden_c256_ui(gen256):
    """The B=1 UI-mode output of the unmodified reference at full size (hair_editor.py:159-179 path)."""
    g = np.load(os.path.join(GOLD, "gen_c256_b1_blocky_ui.npz"))
    labels = torch.from_numpy(g["labels"])
    codes, noise = synth.make_codes(1), synth.make_noise(1, 256)
    out = gen256.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(noise).cuda()).cpu()
    l2, mx = _errs(out, torch.from_numpy(g["out"]))
    assert l2 < L2_TOL and mx < MAX_TOL, (l2, mx)
    # reference-signature entry point: one-hot seg + obj_dic (generator.py:72, normalization.py:121-139)
    seg = so.one_hot(labels)
    obj_dic = {str(j): {"ACE": codes[0, j]} for j in range(19)}
    out2 = gen256(seg, None, obj_dic=obj_dic, noise=synth.flatten_noise(noise)).cpu()
    assert torch.equal(out, out2)


def test_simt_checker_agrees(synthetic_sd, gen64):
    B = 1
    labels, codes, noise = synth.make_labels(B, 64, "iid"), synth.make_codes(B), synth.make_noise(B, 64)
    ref = so.generator_forward(synthetic_sd, labels, codes, noise)
    gen64.impl = _lib.IMPL_SIMT_DEBUG
    try:
        out = gen64.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(noise).cuda()).cpu()
    finally:
        gen64.impl = _lib.IMPL_TCGEN05
    l2, mx = _errs(out, ref)
    assert l2 < L2_TOL and mx < MAX_TOL


def test_host_entry_point_equals_device_entry_point(gen64):
    B = 3
    labels, codes, noise = synth.make_labels(B, 64, "blocky"), synth.make_codes(B), synth.make_noise(B, 64)
    flat = synth.flatten_noise(noise)
    a = gen64.forward_labels(labels.cuda(), codes.cuda(), noise=flat.cuda()).cpu()
    b = gen64.forward_host(labels.numpy(), codes.numpy(), noise=flat.numpy())
    assert torch.equal(a, b)


def test_streamed_host_entry_point_equals_device_entry_point(gen64):
    """forward_host_async double-buffers H2D / compute / D2H across calls: five different batches in flight through
    two buffer sets must each come back equal to the device entry point with the same seed."""
    B, n = 3, 5
    ins = [(synth.make_labels(B, 64, "blocky", seed=40 + i).pin_memory(), synth.make_codes(B, seed=50 + i).pin_memory())
           for i in range(n)]
    outs = [torch.empty((B, 3, 64, 64)).pin_memory() for _ in range(n)]
    for i, (lab, cod) in enumerate(ins):
        gen64.forward_host_async(lab, cod, outs[i], seed=7 + i)
    gen64.host_sync()
    for i, (lab, cod) in enumerate(ins):
        want = gen64.forward_labels(lab.cuda(), cod.cuda(), seed=7 + i).cpu()
        assert torch.equal(outs[i], want), i
    with pytest.raises(_lib.ChbError):
        gen64.forward_host_async(ins[0][0].cuda(), ins[0][1], outs[0])


def test_full_size_batch_properties(gen256):
    """Size-independent properties at 256x256: an image's result does not depend on its batch-mates or its slot,
    and device noise is seed-deterministic."""
    B = 8
    labels, codes = synth.make_labels(B, 256, "blocky").cuda(), synth.make_codes(B).cuda()
    planes = synth.make_noise(B, 256)
    flat = synth.flatten_noise(planes)
    out = gen256.forward_labels(labels, codes, noise=flat.cuda())
    assert torch.isfinite(out).all() and float(out.abs().max()) <= 1.0
    for i in (0, 5):
        flat1 = synth.flatten_noise([p[i:i + 1] for p in planes])
        solo = gen256.forward_labels(labels[i:i + 1], codes[i:i + 1], noise=flat1.cuda())
        # tiles never mix images, but the one-image call runs another schedule: split-K on the 8x8 .. 32x32 convs and the
        # h_0 hi+lo split of interactive batches (csrc/generator.cu split_k_few_tiles, small_batch_h0_split).  The two
        # agree to the fp16 storage rounding of the activations (the batch-of-8 schedule itself is permutation-invariant
        # bitwise, below)
        print("solo vs batch of 8: max |diff| %.3g" % float((solo[0] - out[i]).abs().max()))
        assert float((solo[0] - out[i]).abs().max()) < MAX_TOL
    perm = torch.tensor([3, 1, 7, 0, 2, 6, 5, 4])
    flatp = synth.flatten_noise([p[perm] for p in planes])
    outp = gen256.forward_labels(labels[perm], codes[perm], noise=flatp.cuda())
    assert torch.equal(outp, out[perm])
    a = gen256.forward_labels(labels, codes, seed=11)
    b = gen256.forward_labels(labels, codes, seed=11)
    c = gen256.forward_labels(labels, codes, seed=12)
    assert torch.equal(a, b) and not torch.equal(a, c)


def test_graph_replay_equals_plain_launches(gen64, gen256):
    """chb_generator_forward_graph (one captured CUDA graph per batch size, seed re-parameterised per call) returns
    exactly what the launch-by-launch entry point returns, call after call and for two batch sizes."""
    for g, crop in ((gen64, 64), (gen256, 256)):
        for B in (1, 2):
            labels, codes = synth.make_labels(B, crop, "blocky", seed=3 + B).cuda(), synth.make_codes(B, seed=9).cuda()
            for seed in (5, 6, 5):
                a = g.forward_labels(labels, codes, seed=seed)
                b = g.forward_labels(labels, codes, seed=seed, graph=True)
                assert torch.equal(a, b), (crop, B, seed)
            labels2 = synth.make_labels(B, crop, "iid", seed=77).cuda()   # new inputs through the same graph
            assert torch.equal(g.forward_labels(labels2, codes, seed=8), g.forward_labels(labels2, codes, seed=8, graph=True))


def test_absent_class_codes_ignored(gen64):
    labels = torch.full((1, 64, 64), 3, dtype=torch.uint8)
    labels[:, :32] = 7
    codes = synth.make_codes(1)
    flat = synth.flatten_noise(synth.make_noise(1, 64)).cuda()
    a = gen64.forward_labels(labels.cuda(), codes.cuda(), noise=flat)
    codes2 = codes.clone()
    codes2[:, 5] += 10.0
    b = gen64.forward_labels(labels.cuda(), codes2.cuda(), noise=flat)
    assert torch.equal(a, b)


def test_errors_are_loud(gen64):
    with pytest.raises(_lib.ChbError):
        gen64.forward_labels(torch.zeros((5, 64, 64), dtype=torch.uint8).cuda(), synth.make_codes(5).cuda())
    with pytest.raises(_lib.ChbError):
        gen64.forward_labels(torch.zeros((1, 32, 32), dtype=torch.uint8).cuda(), synth.make_codes(1).cuda())
    with pytest.raises(_lib.ChbError):
        gen64.forward_labels(torch.zeros((1, 64, 64), dtype=torch.uint8), synth.make_codes(1))


def test_generator_512_vs_oracle(synthetic_sd):
    """Config 4's resolution (512x512): one image against the oracle (dense reference algorithm, fp32)."""
    g = SeanGeneratorB200(crop=512, max_batch=1)
    g.load_state_dict(synthetic_sd)
    labels, codes, noise = synth.make_labels(1, 512, "blocky", seed=21), synth.make_codes(1, seed=22), synth.make_noise(1, 512)
    ref = so.generator_forward(synthetic_sd, labels, codes, noise)
    out = g.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(noise).cuda()).cpu()
    assert tuple(out.shape) == (1, 3, 512, 512) and torch.isfinite(out).all()
    l2, mx = _errs(out, ref)
    assert l2 < L2_TOL and mx < MAX_TOL, (l2, mx)


def test_headline_batch64_fresh_inputs_vs_oracle(synthetic_sd):
    """The same schedule on inputs no other test or benchmark uses (iid label maps, new codes and noise): the default
    policy's per-image max-norm averages 8.1e-4 over 24 such images (tools/gpu_maxnorm_distribution.py), 9.5e-4 at worst."""
    B = 64
    g = SeanGeneratorB200(crop=256, max_batch=B)
    g.load_state_dict(synthetic_sd)
    labels, codes = synth.make_labels(B, 256, "iid", seed=102), synth.make_codes(B, seed=152)
    planes = synth.make_noise(B, 256)
    out = g.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(planes).cuda())
    for i in (5, 20, 40, 55):
        ref = so.generator_forward(synthetic_sd, labels[i:i + 1], codes[i:i + 1], [p[i:i + 1] for p in planes])
        l2, mx = _errs(out[i:i + 1].cpu(), ref)
        assert l2 < L2_TOL and mx < MAX_TOL, (i, l2, mx)


def test_headline_batch64_vs_oracle(synthetic_sd):
    """BASELINE.json configs[1] itself: B = 64 at 256x256.  The batch size changes the plan (fc_mu tile width, the N tile
    of the Weff GEMMs, the tile schedule of every persistent launch), so the headline schedule is compared with the
    oracle on the first, a middle and the last image, and image-by-image with the B = 1 schedule (which splits K and keeps
    h_0 as a hi+lo pair: equal to the fp16 storage rounding, not bitwise)."""
    B = 64
    g = SeanGeneratorB200(crop=256, max_batch=B)
    g.load_state_dict(synthetic_sd)
    labels, codes = synth.make_labels(B, 256, "blocky"), synth.make_codes(B)
    planes = synth.make_noise(B, 256)
    out = g.forward_labels(labels.cuda(), codes.cuda(), noise=synth.flatten_noise(planes).cuda())
    assert torch.isfinite(out).all()
    for i in (0, 31, 63):
        one = [p[i:i + 1] for p in planes]
        ref = so.generator_forward(synthetic_sd, labels[i:i + 1], codes[i:i + 1], one)
        l2, mx = _errs(out[i:i + 1].cpu(), ref)
        assert l2 < L2_TOL and mx < MAX_TOL, (i, l2, mx)
        solo = g.forward_labels(labels[i:i + 1].cuda(), codes[i:i + 1].cuda(), noise=synth.flatten_noise(one).cuda())
        print("solo vs batch of 64: max |diff| %.3g" % float((solo[0] - out[i]).abs().max()))
        assert float((solo[0] - out[i]).abs().max()) < MAX_TOL, i
