"""Pins oracle/bisenet_oracle.py against the unmodified reference BiSeNet and writes tests/golden/bisenet_b1.npz.
Run in the build container (needs /root/reference, torchvision and cv2):   python oracle/make_golden_bisenet.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CHB_REFERENCE", "/root/reference")

from ctrlhair_b200 import synth  # noqa: E402
from oracle import bisenet_oracle as bno  # noqa: E402


def main():
    import cv2
    import torch.utils.model_zoo as mz
    import torchvision
    # resnet.py:83 downloads ImageNet weights at construction; they are overwritten by load_state_dict below
    mz.load_url = lambda url, *a, **k: torchvision.models.resnet18(weights=None).state_dict()
    sys.path.insert(0, REF)
    from external_code.face_parsing.model import BiSeNet   # the unmodified reference

    sd = synth.make_bisenet_state_dict()
    net = BiSeNet(n_classes=19).eval()
    assert list(net.state_dict().keys()) == list(sd.keys()), "synthetic checkpoint keys differ from the reference's"
    for k, v in net.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
    net.load_state_dict(sd, strict=True)

    g = np.random.default_rng(5)
    yy, xx = np.mgrid[0:512, 0:512]
    img = (128 + 80 * np.sin(xx / 37.0)[..., None] * np.cos(yy / 53.0)[..., None] + g.normal(0, 25, (512, 512, 3)))
    img = np.clip(img, 0, 255).astype(np.uint8)[None]
    x = bno.normalise_image(img)
    # my_parsing_util.py:25-36 on the same (already 512x512) image
    import torchvision.transforms as transforms
    to_tensor = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
    from PIL import Image
    x_ref = to_tensor(Image.fromarray(img[0]))[None]
    assert float((x - x_ref).abs().max()) < 1e-6
    with torch.no_grad():
        want = net(x_ref)[0]
        got = bno.bisenet_forward(sd, x)
    err = float((got - want).abs().max() / want.abs().max())
    print("bisenet logits: max-norm error of the restatement %.2e, logit range %.2f .. %.2f, classes used %d" %
          (err, float(want.min()), float(want.max()), len(np.unique(want.argmax(1).numpy()))))
    assert err < 1e-5
    parsing = want.squeeze(0).cpu().numpy().argmax(0)                       # my_parsing_util.py:45-46
    assert np.array_equal(parsing, bno.parsing_labels(got)[0])
    # my_parsing_util.py:49-54 executed literally
    sys.path.insert(0, REF)
    from global_value_utils import PARSING_LABEL_LIST
    assert PARSING_LABEL_LIST == bno.PARSING_LABEL_LIST
    label_lists = bno.BISENET_LABELS
    celeba = np.zeros_like(parsing)
    for label_idx, label_name in enumerate(PARSING_LABEL_LIST):
        celeba[label_lists.index(label_name) == parsing] = label_idx
    assert np.array_equal(celeba, bno.swap_parsing_label_to_celeba_mask(parsing))
    mask = cv2.resize(celeba.astype("uint8"), (256, 256), interpolation=cv2.INTER_NEAREST)   # hair_editor.py:334
    assert np.array_equal(mask, bno.get_mask(sd, img)[0])
    path = os.path.join(ROOT, "tests", "golden", "bisenet_b1.npz")
    np.savez_compressed(path, img=img, logits_sub=want[:, :, ::16, ::16].numpy(), parsing=parsing.astype(np.uint8),
                        mask256=mask)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
