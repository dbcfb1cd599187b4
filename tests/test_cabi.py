"""The C-ABI library loads on a CPU-only box and exports every symbol include/ctrlhair_b200.h declares."""
import ctypes as C
import os
import re

import torch

from ctrlhair_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctrlhair_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(chb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    bound = {s[0] for s in _lib.SYMBOLS}
    assert bound == set(names), "ctypes binding and header disagree: %s" % (bound ^ set(names))


def test_struct_sizes_match_header(lib):
    # guards against silent ABI drift of the ctypes mirror: the library reports sizeof() of its own structs
    assert C.sizeof(_lib.ConvSeg) == lib.chb_struct_size(0) == 80
    assert C.sizeof(_lib.ConvDesc) == lib.chb_struct_size(1)
    assert C.sizeof(_lib.GenConfig) == lib.chb_struct_size(2) == 24
    assert C.sizeof(_lib.MlpLayer) == lib.chb_struct_size(3)
    assert lib.chb_struct_size(99) == -1


def test_version_and_argument_errors(lib):
    assert lib.chb_version() == 100
    assert lib.chb_conv_run(None, 0, None) == -1
    assert b"NULL" in lib.chb_last_error()
    cfg = _lib.GenConfig(48, 19, 256, 512, 1)  # ngf not a multiple of 64
    h = C.c_void_p()
    assert lib.chb_generator_create(C.byref(cfg), C.byref(h)) == -1
    cfg = _lib.GenConfig(64, 19, 250, 512, 1)  # crop not a power of two
    assert lib.chb_generator_create(C.byref(cfg), C.byref(h)) == -1


def test_no_cpu_fallback():
    """Without a GPU the product path must fail loudly, not compute on the CPU."""
    if torch.cuda.is_available():
        return
    from ctrlhair_b200.generator import SeanGeneratorB200
    from ctrlhair_b200 import ops
    try:
        SeanGeneratorB200()
        raise AssertionError("expected ChbError")
    except _lib.ChbError:
        pass
    try:
        ops.onehot_pyramid(torch.zeros((1, 32, 32), dtype=torch.uint8), [32])
        raise AssertionError("expected ChbError")
    except _lib.ChbError:
        pass
    lib = _lib.load()
    assert lib.chb_check_device() != 0


def test_postprocessing_entry_points_validate_arguments(lib):
    """Argument errors of the post-processing entry points are reported before anything touches the device."""
    assert lib.chb_postprocess_workspace_bytes(2, 256, 256) == 2 * 256 * 256 * 4
    assert lib.chb_postprocess_workspace_bytes(0, 256, 256) == 0
    assert lib.chb_image_to_u8(None, None, 1, 8, 8, None) == -1
    assert lib.chb_blend_mask(None, None, None, None, 1, 8, 8, None) == -1
    assert lib.chb_rgb_to_hsv(None, None, None, 1, None) == -1
    assert b"exactly one" in lib.chb_last_error()
    assert lib.chb_hsv_to_rgb(None, None, 1, None) == -1
    assert lib.chb_onehot_to_label(None, None, 1, 19, 64, None) == -1
    assert lib.chb_label_to_onehot(None, None, 1, 19, 64, None) == -1
    dummy = C.c_void_p(16)   # never dereferenced: the size checks come first
    assert lib.chb_poisson_blend(dummy, dummy, dummy, dummy, 1, 300, 300, 1, 1e-11, 100, None, None, None, None) == -1
    assert b"too large" in lib.chb_last_error()
    assert lib.chb_poisson_blend(dummy, dummy, dummy, dummy, 1, 2, 8, 1, 1e-11, 100, None, None, None, None) == -1
    assert lib.chb_poisson_blend(dummy, dummy, dummy, dummy, 1, 64, 64, 1, 0.0, 100, None, None, None, None) == -1
    assert lib.chb_postprocess_blending(None, None, None, None, None, None, None, 1, 8, 8, 1, 1e-11, 10, None, None,
                                        None, None) == -1
