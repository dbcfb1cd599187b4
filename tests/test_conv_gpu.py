"""Parity of the tcgen05 implicit-GEMM conv operator (through the C ABI) against a plain torch fp32 reference."""
import pytest
import torch

import _conv_cases as cc
from ctrlhair_b200 import ops

pytestmark = pytest.mark.gpu

# fp16 operands, fp32 accumulate; outputs stored in fp16 carry one more rounding (2^-11 relative)
TOL = 1e-3


@pytest.fixture(scope="module")
def cases():
    assert torch.cuda.is_available()
    return dict(cc.make_cases())


@pytest.mark.parametrize("name", [n for n, _ in cc.make_cases(device="cpu")])
def test_conv_case_tcgen05(cases, name):
    got, want = cases[name](ops.IMPL_TCGEN05)
    torch.cuda.synchronize()
    assert torch.isfinite(got).all()
    assert cc.rel_err(got, want) < TOL


@pytest.mark.parametrize("name", ["plain_conv1_plus_convs", "mod_styled_c128_bn256_up", "plain_tb4_rows"])
def test_conv_case_simt_checker(cases, name):
    got, want = cases[name](ops.IMPL_SIMT_DEBUG)
    assert cc.rel_err(got, want) < TOL


def test_split_k_is_deterministic_and_close_to_unsplit():
    """The CTA that arrives last adds the partial accumulators in split order, so repeated launches agree bitwise; against
    the unsplit launch only the fp32 association differs."""
    gen = torch.Generator().manual_seed(5)
    a = (torch.randn((2, 8, 8, 1024), generator=gen)).to("cuda", torch.float16)
    w = (torch.randn((256, 9 * 1024), generator=gen) / 96).to("cuda", torch.float16)
    bias = torch.randn((256,), generator=gen).cuda()
    seg = [dict(a=a, w=w, C=1024, taps=9)]
    base = ops.conv_igemm(seg, 256, 64, bias=bias, out_dtype=torch.float32, tile=(8, 8, 2))
    runs = [ops.conv_igemm(seg, 256, 64, bias=bias, out_dtype=torch.float32, tile=(8, 8, 2), ksplit=4) for _ in range(5)]
    for r in runs[1:]:
        assert torch.equal(r, runs[0])
    assert float((runs[0] - base).abs().max()) < 2e-5 * float(base.abs().max())


def test_conv_rejects_bad_descriptors():
    from ctrlhair_b200 import _lib
    a = torch.zeros((1, 8, 8, 48), dtype=torch.float16, device="cuda")  # 48 channels: not 32 / multiple of 64
    w = torch.zeros((64, 9 * 48), dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.ChbError):
        ops.conv_igemm([dict(a=a, w=w)], 64, 64)
    with pytest.raises(_lib.ChbError):  # host tensor: there is no CPU path
        ops.conv_igemm([dict(a=a.cpu(), w=w)], 64, 64)
    a64 = torch.zeros((1, 8, 8, 64), dtype=torch.float16, device="cuda")
    w64 = torch.zeros((64, 9 * 64), dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.ChbError):  # more splits than the segment has channel chunks
        ops.conv_igemm([dict(a=a64, w=w64)], 64, 64, ksplit=2)
