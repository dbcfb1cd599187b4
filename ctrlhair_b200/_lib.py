"""ctypes binding of include/ctrlhair_b200.h.  The library is built in-tree by ctrlhair_b200/build.py.

There is deliberately no fallback: if the shared library is missing or the device is not sm_100, every
entry point raises.
"""
import ctypes as C
import os

from . import build as _build

_LIB = None


class ChbError(RuntimeError):
    pass


class ConvSeg(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("a_sb", C.c_int64), ("a_sy", C.c_int64), ("a_sx", C.c_int64),
        ("Ca", C.c_int), ("ch_off", C.c_int), ("C", C.c_int), ("taps", C.c_int),
        ("w", C.c_void_p),
        ("per_image", C.c_int),
        ("w_sb", C.c_int64),
        ("a_pad", C.c_int),
        ("w_dup", C.c_int),
    ]


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("TW", C.c_int), ("TH", C.c_int), ("TB", C.c_int),
        ("nseg", C.c_int),
        ("seg", ConvSeg * 4),
        ("N", C.c_int), ("Nrows", C.c_int), ("BN", C.c_int),
        ("epi", C.c_int), ("act", C.c_int),
        ("bias", C.c_void_p), ("bias_per_image", C.c_int),
        ("out", C.c_void_p), ("out_dtype", C.c_int),
        ("o_sb", C.c_int64), ("o_sy", C.c_int64), ("o_sx", C.c_int64), ("o_sn", C.c_int64),
        ("o_ngroup", C.c_int), ("o_sgroup", C.c_int64),
        ("res", C.c_void_p), ("r_sb", C.c_int64), ("r_sy", C.c_int64), ("r_sx", C.c_int64), ("r_shift", C.c_int),
        ("x", C.c_void_p), ("x_sb", C.c_int64), ("x_sy", C.c_int64), ("x_sx", C.c_int64), ("x_shift", C.c_int),
        ("noise", C.c_void_p),
        ("chan", C.c_void_p),
        ("o_split", C.c_int), ("o_lo_off", C.c_int64),
        ("ksplit", C.c_int), ("ks_ws", C.c_void_p),
    ]


class MlpLayer(C.Structure):
    _fields_ = [("in_dim", C.c_int), ("out_dim", C.c_int), ("wt", C.c_void_p), ("bias", C.c_void_p),
                ("pre_act", C.c_int), ("post_act", C.c_int), ("inj_u", C.c_void_p), ("inj_l", C.c_void_p),
                ("inj_mu", C.c_void_p), ("inj_nb", C.c_int), ("inj_zoff", C.c_int)]


class GenConfig(C.Structure):
    _fields_ = [("ngf", C.c_int), ("label_nc", C.c_int), ("crop", C.c_int), ("style_len", C.c_int),
                ("max_batch", C.c_int), ("precision", C.c_uint)]


PREC_IMG, PREC_SHORTCUT = 1, 2


def prec_h1(i):
    return 1 << (8 + i)


def prec_h0(i):
    return 1 << (16 + i)


def prec_w(i):
    return 1 << (24 + i)


class ZencConfig(C.Structure):
    _fields_ = [("crop", C.c_int), ("label_nc", C.c_int), ("max_batch", C.c_int)]


class BisenetConfig(C.Structure):
    _fields_ = [("size", C.c_int), ("n_classes", C.c_int), ("max_batch", C.c_int)]


class ShapeConfig(C.Structure):
    _fields_ = [("crop", C.c_int), ("max_batch", C.c_int)]


class CtTrainConfig(C.Structure):
    _fields_ = [("batch", C.c_int)] + [(k, C.c_float) for k in (
        "lambda_adv", "lambda_gp", "lambda_info", "lambda_info_curliness", "lambda_rec", "lambda_rgb", "lambda_pca_std",
        "lambda_moment_1", "lambda_moment_2", "lambda_cls_curliness", "lambda_orthogonal", "lr", "beta1", "beta2",
        "eps")] + [("use_graph", C.c_int)]


class CtTrainBatch(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in (
        "code", "rgb_mean", "pca_std", "noise", "noise_curliness", "curliness_label", "perm_rgb", "perm_curliness",
        "perm_noise", "alpha_gp")] + [("noise_from_encoder", C.c_int)]


CTT_D, CTT_G, CTT_FROZEN = 0, 1, 2
CTT_LOSS_NAMES = ("lambda_adv", "lambda_gp", "lambda_info", "lambda_rec", "lambda_moment_1", "lambda_moment_2",
                  "lambda_info_curliness", "lambda_rgb", "lambda_pca_std", "lambda_cls_curliness", "lambda_orthogonal",
                  "total")
EPI_PLAIN, EPI_MODULATE = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
F16, F32 = 0, 1
IMPL_TCGEN05, IMPL_SIMT_DEBUG = 0, 1

# every symbol include/ctrlhair_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("chb_version", C.c_int, []),
    ("chb_last_error", C.c_char_p, []),
    ("chb_check_device", C.c_int, []),
    ("chb_conv_ksplit_workspace_bytes", C.c_int64, [C.c_int]),
    ("chb_conv_run", C.c_int, [C.POINTER(ConvDesc), C.c_int, C.c_void_p]),
    ("chb_onehot_pyramid", C.c_int,
     [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int, C.c_void_p]),
    ("chb_noise_fill", C.c_int, [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]),
    ("chb_f32_to_f16", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("chb_mlp_forward", C.c_int,
     [C.POINTER(MlpLayer), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_struct_size", C.c_int, [C.c_int]),
    ("chb_generator_create", C.c_int, [C.POINTER(GenConfig), C.POINTER(C.c_void_p)]),
    ("chb_generator_destroy", None, [C.c_void_p]),
    ("chb_generator_num_tensors", C.c_int, [C.c_void_p]),
    ("chb_generator_tensor_info", C.c_int,
     [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    ("chb_generator_blob_bytes", C.c_int64, [C.c_void_p]),
    ("chb_generator_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("chb_generator_bind", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_generator_forward", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("chb_generator_forward_graph", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_generator_forward_host", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("chb_generator_forward_host_async", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_generator_host_sync", C.c_int, [C.c_void_p]),
    ("chb_generator_forward_timed", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_void_p,
      C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int]),
    ("chb_generator_step_name", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_int]),
    ("chb_generator_noise_floats", C.c_int64, [C.c_void_p, C.c_int]),
    ("chb_generator_launches", C.c_int, [C.c_void_p]),
    ("chb_generator_flops", C.c_double, [C.c_void_p, C.c_int]),
    ("chb_generator_set_step_limit", C.c_int, [C.c_void_p, C.c_int]),
    ("chb_generator_debug_tensor", C.c_int64,
     [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int)]),
    ("chb_bisenet_create", C.c_int, [C.POINTER(BisenetConfig), C.POINTER(C.c_void_p)]),
    ("chb_bisenet_destroy", None, [C.c_void_p]),
    ("chb_bisenet_num_tensors", C.c_int, [C.c_void_p]),
    ("chb_bisenet_tensor_info", C.c_int,
     [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    ("chb_bisenet_blob_bytes", C.c_int64, [C.c_void_p]),
    ("chb_bisenet_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("chb_bisenet_launches", C.c_int, [C.c_void_p]),
    ("chb_bisenet_bind", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_bisenet_forward", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_bisenet_forward_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("chb_pil_resize_bilinear", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
      C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_zencoder_create", C.c_int, [C.POINTER(ZencConfig), C.POINTER(C.c_void_p)]),
    ("chb_zencoder_destroy", None, [C.c_void_p]),
    ("chb_zencoder_num_tensors", C.c_int, [C.c_void_p]),
    ("chb_zencoder_tensor_info", C.c_int,
     [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    ("chb_zencoder_blob_bytes", C.c_int64, [C.c_void_p]),
    ("chb_zencoder_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("chb_zencoder_bind", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_zencoder_forward", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_zencoder_forward_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_shape_create", C.c_int, [C.POINTER(ShapeConfig), C.POINTER(C.c_void_p)]),
    ("chb_shape_destroy", None, [C.c_void_p]),
    ("chb_shape_num_tensors", C.c_int, [C.c_void_p]),
    ("chb_shape_tensor_info", C.c_int,
     [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    ("chb_shape_blob_bytes", C.c_int64, [C.c_void_p]),
    ("chb_shape_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("chb_shape_bind", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_shape_encode", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_shape_encode_labels", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_shape_decode", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_shape_decode_labels", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_shape_decode_logits", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_shape_softmax", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_cttrain_create", C.c_int, [C.POINTER(CtTrainConfig), C.POINTER(C.c_void_p)]),
    ("chb_cttrain_destroy", None, [C.c_void_p]),
    ("chb_cttrain_num_tensors", C.c_int, [C.c_void_p]),
    ("chb_cttrain_tensor_info", C.c_int,
     [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    ("chb_cttrain_state_floats", C.c_int64, [C.c_void_p]),
    ("chb_cttrain_region", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("chb_cttrain_workspace_bytes", C.c_int64, [C.c_void_p]),
    ("chb_cttrain_bind", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_cttrain_step", C.c_int, [C.c_void_p, C.c_int, C.POINTER(CtTrainBatch), C.c_void_p, C.c_void_p]),
    ("chb_cttrain_adam", C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    ("chb_cttrain_launches", C.c_int, [C.c_void_p, C.c_int]),
    ("chb_cttrain_schedule", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("chb_image_to_u8", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("chb_blend_mask", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("chb_poisson_blend", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_poisson_coarse_inverse", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("chb_postprocess_workspace_bytes", C.c_int64, [C.c_int, C.c_int, C.c_int]),
    ("chb_postprocess_blending", C.c_int,
     [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
      C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("chb_rgb_to_hsv", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("chb_hsv_to_rgb", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("chb_onehot_to_label", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    ("chb_label_to_onehot", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
]


def load(build_if_missing=True):
    """Loads (building first if needed) the CUDA library.  Raises if it cannot be had."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("CHB_LIB_PATH") or _build.lib_path()  # override: A/B runs of two builds on one box
    if not os.path.exists(path):
        if not build_if_missing:
            raise ChbError("ctrlhair_b200: %s is missing (run __graft_entry__.build())" % path)
        _build.build_library()
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    for which, mirror in ((0, ConvSeg), (1, ConvDesc), (2, GenConfig), (3, MlpLayer)):
        if lib.chb_struct_size(which) != C.sizeof(mirror):
            raise ChbError("ctrlhair_b200: %s was built from a different include/ctrlhair_b200.h than this binding "
                           "(struct %s: %d vs %d bytes); rebuild with __graft_entry__.build()" %
                           (path, mirror.__name__, lib.chb_struct_size(which), C.sizeof(mirror)))
    _LIB = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().chb_last_error()
        raise ChbError("ctrlhair_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
