"""TEST INFRASTRUCTURE — runs only in the build container, where /root/reference exists.

Imports the *unmodified* reference (XuyangGuo/CtrlHair) from /root/reference with the minimum shims needed to
run it on CPU, so that oracle/make_golden.py can record golden input/output vectors and validate the oracle
restatement (oracle/sean_oracle.py).  Nothing in tests -m gpu, smoke() or bench.py imports this module.

Shims (all outside the reference tree):
  * torch.Tensor.cuda -> identity          (normalization.py:111 and hair_editor.py:146 hard-code .cuda())
  * torch.randn       -> injected planes   (normalization.py:111 draws one plane per ACE call)
  * opt               -> argparse.Namespace with the values of sean_codes/options/base_options.py
"""
import argparse
import contextlib
import os
import sys

import torch

REF_ROOT = os.environ.get("CTRLHAIR_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "sean_codes"))


def make_opt(ngf=64, crop=256, label_nc=19):
    # values: sean_codes/options/base_options.py:19-72, generator.py:16-22
    return argparse.Namespace(
        ngf=ngf, label_nc=label_nc, semantic_nc=label_nc, crop_size=crop, aspect_ratio=1.0,
        num_upsampling_layers="normal", norm_G="spectralspadesyncbatch3x3", status="test", gpu_ids=[],
        contain_dontcare_label=False, no_instance=True, init_type="xavier", init_variance=0.02, isTrain=False,
        use_vae=False, netG="spade")


@contextlib.contextmanager
def reference_on_path():
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    old_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, REF_ROOT)
    try:
        yield
    finally:
        sys.path.remove(REF_ROOT)
        torch.Tensor.cuda = old_cuda


@contextlib.contextmanager
def injected_randn(planes):
    """Replaces torch.randn by a queue of pre-drawn planes (one per ACE call, call order)."""
    queue = list(planes)
    real = torch.randn

    def fake(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        t = queue.pop(0)
        assert tuple(t.shape) == tuple(shape), (t.shape, shape)
        return t.clone()

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = real
        assert not queue, "reference drew fewer noise planes than provided"


def build_reference_generator(state_dict, ngf=64, crop=256):
    """Constructs the reference SPADEGenerator (generator.py:14-53) and loads `state_dict` strictly."""
    with reference_on_path():
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from sean_codes.models.networks.generator import SPADEGenerator
        net = SPADEGenerator(make_opt(ngf, crop))
    net.load_state_dict(state_dict, strict=True)
    net.eval()
    return net


def set_status(net, status):
    # hair_editor.py:34-37
    for m in net.modules():
        if hasattr(m, "status"):
            m.status = status


def run_reference_generator(net, onehot, codes, noise_planes):
    """Batched reference forward: replicates SPADEGenerator.forward (generator.py:72-109) with
    style_codes given directly (status='train' so ACE takes the per-image branch, normalization.py:141-153;
    SURVEY A8: identical to the UI path looped at B=1)."""
    import torch.nn.functional as F
    set_status(net, "train")
    with reference_on_path(), injected_randn(noise_planes), torch.no_grad():
        seg = onehot
        x = F.interpolate(seg, size=(net.sh, net.sw))
        x = net.fc(x)
        x = net.head_0(x, seg, codes, obj_dic=None)
        x = net.up(x)
        x = net.G_middle_0(x, seg, codes, obj_dic=None)
        x = net.G_middle_1(x, seg, codes, obj_dic=None)
        x = net.up(x)
        x = net.up_0(x, seg, codes, obj_dic=None)
        x = net.up(x)
        x = net.up_1(x, seg, codes, obj_dic=None)
        x = net.up(x)
        x = net.up_2(x, seg, codes, obj_dic=None)
        x = net.up(x)
        x = net.up_3(x, seg, codes, obj_dic=None)
        x = net.conv_img(F.leaky_relu(x, 2e-1))
        return torch.tanh(x)


def run_reference_ui(net, onehot, codes_1, noise_planes):
    """The UI path exactly as hair_editor.py:159-179 drives it: B=1, obj_dic, empty rgb image."""
    set_status(net, "UI_mode")
    obj_dic = {str(j): {"ACE": codes_1[j]} for j in range(codes_1.shape[0])}
    empty = torch.zeros((0, 3, onehot.shape[2], onehot.shape[3]))
    with reference_on_path(), injected_randn(noise_planes), torch.no_grad():
        return net(onehot, empty, obj_dic=obj_dic)
