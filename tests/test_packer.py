"""Host logic: the weight packer's folds/layouts and the region-factored schedule, checked on CPU."""
import ctypes as C

import pytest
import torch

import _packed_emulation as pe
from ctrlhair_b200 import _lib, packer
from oracle import sean_oracle as so
from ctrlhair_b200 import synth


def test_tile_gamma_beta_roundtrip():
    g = torch.arange(512.).reshape(512, 1)
    b = -torch.arange(512.).reshape(512, 1)
    t = packer.tile_gamma_beta(g, b, 256)
    assert t.shape == (1024, 1)
    assert torch.equal(t[:128, 0], g[:128, 0]) and torch.equal(t[128:256, 0], b[:128, 0])
    assert torch.equal(t[256:384, 0], g[128:256, 0])
    g2, b2 = pe._untile(t.reshape(1, 1024, 1, 1), 256)
    assert torch.equal(g2.reshape(-1), g.reshape(-1)) and torch.equal(b2.reshape(-1), b.reshape(-1))


@pytest.mark.parametrize("kind", ["blocky", "iid"])
def test_factored_schedule_is_exact_in_fp32(synthetic_sd, kind):
    """Packed + factored form == dense reference form, to fp32 round-off."""
    labels, codes, noise = synth.make_labels(2, 64, kind), synth.make_codes(2), synth.make_noise(2, 64)
    ref = so.generator_forward(synthetic_sd, labels, codes, noise)
    packed = packer.pack_generator(synthetic_sd, weight_dtype=torch.float32)
    out = pe.emulate(packed, labels, codes, noise)
    assert float((out - ref).abs().max() / ref.abs().max()) < 2e-5


def test_fp16_storage_forecast(synthetic_sd):
    """fp16 weights + fp16 MMA operands, fp32 everywhere else: the error budget the CUDA path is held to."""
    labels, codes, noise = synth.make_labels(2, 64, "iid"), synth.make_codes(2), synth.make_noise(2, 64)
    ref = so.generator_forward(synthetic_sd, labels, codes, noise)
    packed = packer.pack_generator(synthetic_sd)
    out = pe.emulate(packed, labels, codes, noise, round16=True)
    assert float((out - ref).norm() / ref.norm()) < 1e-3
    assert float((out - ref).abs().max() / ref.abs().max()) < 3e-3
    # with the hi+lo split policies (chb_gen_config.precision) the forecast falls under north_star's 1e-3 max-norm
    from ctrlhair_b200.generator import PRECISION_POLICIES
    errs = {}
    for policy in ("shortcut", "parity", "full"):
        o = pe.emulate(packed, labels, codes, noise, round16=True, precision=PRECISION_POLICIES[policy])
        errs[policy] = float((o - ref).abs().max() / ref.abs().max())
    assert errs["parity"] < 1e-3 and errs["full"] < 1e-3 and errs["full"] < errs["shortcut"], errs


@pytest.mark.parametrize("policy", ["fast", "shortcut", "h1", "parity", "full", "margin"])
def test_packer_matches_library_layout(synthetic_sd, lib, policy):
    """Every tensor the library's blob layout names is produced by the packer with the right dtype and size."""
    from ctrlhair_b200.generator import PRECISION_POLICIES
    cfg = _lib.GenConfig(64, 19, 256, 512, 4, PRECISION_POLICIES[policy])
    h = C.c_void_p()
    assert lib.chb_generator_create(C.byref(cfg), C.byref(h)) == 0
    try:
        packed = packer.pack_generator(synthetic_sd)
        n = lib.chb_generator_num_tensors(h)
        name = C.create_string_buffer(128)
        off, nb, dt = C.c_int64(), C.c_int64(), C.c_int()
        seen, end = set(), 0
        for i in range(n):
            assert lib.chb_generator_tensor_info(h, i, name, 128, C.byref(off), C.byref(nb), C.byref(dt)) == 0
            k = name.value.decode()
            t = packed[k]
            assert t.dtype == (torch.float16 if dt.value == _lib.F16 else torch.float32), k
            assert t.numel() * t.element_size() == nb.value, k
            assert off.value % 256 == 0 and off.value >= end
            end = off.value + nb.value
            seen.add(k)
        flags = PRECISION_POLICIES[policy]
        blocks = [b[0] for b in packer.BLOCKS]

        def placed(k):   # which optional tensors this policy puts into the blob
            if not packer.is_optional(k):
                return True
            if k.endswith(".conv_s.wlo"):
                return bool(flags & _lib.PREC_SHORTCUT)
            if k.startswith("conv_img."):
                return bool(flags & _lib.PREC_IMG)
            return bool(flags & _lib.prec_w(blocks.index(k.split(".")[0])))
        assert seen == {k for k in packed if placed(k)}
        assert lib.chb_generator_blob_bytes(h) >= end
        # 267 M reference parameters -> ~534 MB of fp16 (one-hot padding 19->32 adds a little)
        assert 5.0e8 < lib.chb_generator_blob_bytes(h) < 6.0e8
        # 3 helpers + fc_mu + 4 grouped weff + fc + 7 blocks x (mlp_shared + 2-3 ACEs + 2 convs) + conv_img (+ gather)
        assert lib.chb_generator_launches(h) == (49 if policy == "fast" else 50)
    finally:
        lib.chb_generator_destroy(h)
