"""Runs the UNMODIFIED reference SPADEGenerator (staged by baseline/make_ref.py) for bench.py's reference arm.

The reference's own call path is timed: `SPADEGenerator.forward` in `.eval()`, status 'UI_mode', one image per call
with an `obj_dic` of per-class style codes — what `HairEditor.gen_img` does (hair_editor.py:159-179 ->
pix2pix_model.py:59-68,119-144,208-215 -> generator.py:72-109).  Nothing of ctrlhair_b200's kernels or packer is on
this path; `ctrlhair_b200.synth` only supplies the reference-format synthetic checkpoint and the inputs.

Shims, all outside the reference tree:
  * CPU runs only: torch.Tensor.cuda -> identity (normalization.py:111 hard-codes `.cuda()`); on the GPU nothing is patched
  * `opt` is an argparse.Namespace with the values of sean_codes/options/base_options.py (TestOptions().parse() reads sys.argv)
"""
import argparse
import contextlib
import json
import os
import sys
import time
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")


def ref_root():
    """Staged copy first (it is what exists on the GPU box), else the container's read-only reference tree."""
    if os.path.isdir(os.path.join(STAGED, "sean_codes", "models", "networks")):
        return STAGED
    alt = os.environ.get("CTRLHAIR_REFERENCE", "/root/reference")
    if os.path.isdir(os.path.join(alt, "sean_codes", "models", "networks")):
        return alt
    return None


def ref_digest():
    try:
        return json.load(open(os.path.join(STAGED, "DIGEST.json")))["sha256"][:16]
    except Exception:
        return None


def make_opt(ngf=64, crop=256, label_nc=19):
    # sean_codes/options/base_options.py:19-72, generator.py:16-22
    return argparse.Namespace(
        ngf=ngf, label_nc=label_nc, semantic_nc=label_nc, crop_size=crop, aspect_ratio=1.0,
        num_upsampling_layers="normal", norm_G="spectralspadesyncbatch3x3", status="test", gpu_ids=[],
        contain_dontcare_label=False, no_instance=True, init_type="xavier", init_variance=0.02, isTrain=False,
        use_vae=False, netG="spade")


@contextlib.contextmanager
def _on_path(root, cpu):
    old_cuda = torch.Tensor.cuda
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
    sys.path.insert(0, root)
    try:
        yield
    finally:
        sys.path.remove(root)
        torch.Tensor.cuda = old_cuda


def build_generator(state_dict, crop=256, device="cpu"):
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree not staged (run python baseline/make_ref.py where /root/reference exists)")
    cpu = torch.device(device).type == "cpu"
    with _on_path(root, cpu), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from sean_codes.models.networks.generator import SPADEGenerator
        net = SPADEGenerator(make_opt(crop=crop))
    net.load_state_dict(state_dict, strict=True)
    net.eval()
    for m in net.modules():  # hair_editor.py:34-37 change_status(model, 'UI_mode')
        if hasattr(m, "status"):
            m.status = "UI_mode"
    return net.to(device)


@contextlib.contextmanager
def _injected_randn(planes, device):
    """Parity runs only: torch.randn hands out pre-drawn ACE noise planes (normalization.py:111 draws one per ACE call)."""
    if planes is None:
        yield
        return
    queue = [p.to(device) for p in planes]
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


def forward_ui(net, label_map, codes_1, device, noise_planes=None):
    """One `gen_img`-style call: label map uint8 [H,W] + codes [19,512] -> image [1,3,H,W] (on `device`).
    noise_planes: optional 18 tensors [1,r,r,1] handed to the reference's torch.randn calls (parity checks)."""
    root = ref_root()
    dev = torch.device(device)
    with _on_path(root, dev.type == "cpu"), _injected_randn(noise_planes, dev), torch.no_grad():
        lab = label_map.to(dev).long()[None, None]
        # pix2pix_model.py:131-135: one-hot label map by scatter_
        onehot = torch.zeros((1, 19, lab.shape[2], lab.shape[3]), dtype=torch.float32, device=dev).scatter_(1, lab, 1.0)
        obj_dic = {str(j): {"ACE": codes_1[j].to(dev)} for j in range(codes_1.shape[0])}
        empty = torch.zeros((0, 3, lab.shape[2], lab.shape[3]), device=dev)  # hair_editor.py:170-174
        return net(onehot, empty, obj_dic=obj_dic)


def time_generator(state_dict, labels, codes, crop, device, steps, warmup):
    """images/s of the reference UI path, B = 1 per call, `steps` timed calls after `warmup` untimed ones."""
    dev = torch.device(device)
    net = build_generator(state_dict, crop, dev)
    n = labels.shape[0]
    sync = (lambda: torch.cuda.synchronize(dev)) if dev.type == "cuda" else (lambda: None)
    for i in range(warmup):
        forward_ui(net, labels[i % n], codes[i % n], dev)
    sync()
    t0 = time.perf_counter()
    out = None
    for i in range(warmup, warmup + steps):
        out = forward_ui(net, labels[i % n], codes[i % n], dev)
    sync()
    dt = time.perf_counter() - t0
    finite = bool(torch.isfinite(out).all()) if out is not None else False
    del net
    return steps / dt, dt / max(steps, 1), finite
