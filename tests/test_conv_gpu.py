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


def test_conv_rejects_bad_descriptors():
    from ctrlhair_b200 import _lib
    a = torch.zeros((1, 8, 8, 48), dtype=torch.float16, device="cuda")  # 48 channels: not 32 / multiple of 64
    w = torch.zeros((64, 9 * 48), dtype=torch.float16, device="cuda")
    with pytest.raises(_lib.ChbError):
        ops.conv_igemm([dict(a=a, w=w)], 64, 64)
    with pytest.raises(_lib.ChbError):  # host tensor: there is no CPU path
        ops.conv_igemm([dict(a=a.cpu(), w=w)], 64, 64)
