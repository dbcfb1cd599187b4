"""The reference-facing facade (Pix2PixModel.forward modes, HairEditor.gen_img / get_code) on the GPU path."""
import pytest
import torch

from ctrlhair_b200 import synth
from oracle import sean_oracle as so
from oracle import zencoder_oracle as zo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def editor(synthetic_sd):
    from ctrlhair_b200.editor import HairEditorB200
    gen = torch.Generator().manual_seed(77)
    median = torch.randn((19, 512), generator=gen) * 0.135
    return HairEditorB200(synthetic_sd, median_codes=median, img_size=64), median


def test_gen_img_uses_median_for_zero_rows(synthetic_sd, editor):
    ed, median = editor
    labels = synth.make_labels(1, 64, "blocky")
    code = synth.make_codes(1)
    code[0, 4] = 0  # an all-zero row means "use the median default" (hair_editor.py:165-168)
    code[0, 13] = 0
    noise = synth.make_noise(1, 64)
    img = ed.gen_img(code, labels[:, None].numpy(), noise=synth.flatten_noise(noise)).cpu()
    eff = code.clone()
    eff[0, 4], eff[0, 13] = median[4], median[13]
    ref = so.generator_forward(synthetic_sd, labels, eff, noise)[0]
    assert img.shape == (3, 64, 64)
    assert float((img - ref).norm() / ref.norm()) < 1e-3


def test_style_code_mode_and_invalid_mode(synthetic_sd, editor):
    ed, _ = editor
    img = synth.make_image(1, 64)
    labels = synth.make_labels(1, 64, "blocky")
    codes = ed.get_code(img.numpy(), labels[:, None].numpy()).cpu()
    ref = zo.zencoder_forward(synthetic_sd, img, labels)
    assert codes.shape == (1, 19, 512)
    assert float((codes - ref).norm() / ref.norm()) < 1e-3
    with pytest.raises(ValueError):
        ed.sean_model({"label": labels[:, None].float(), "image": img}, mode="generator")


def test_gen_img_batch_equals_looped_gen_img(synthetic_sd):
    """one batched call == B separate gen_img calls (same noise), zero rows -> median."""
    from ctrlhair_b200.editor import HairEditorB200
    gen = torch.Generator().manual_seed(78)
    median = torch.randn((19, 512), generator=gen) * 0.135
    B = 3
    ed = HairEditorB200(synthetic_sd, median_codes=median, img_size=64, max_batch=B)
    labels = synth.make_labels(B, 64, "blocky", seed=5)
    codes = synth.make_codes(B)
    codes[1, 13] = 0
    codes[2, 0] = 0
    noise = synth.make_noise(B, 64)
    batch = ed.gen_img_batch(codes, labels, noise=synth.flatten_noise(noise).cuda()).cpu()
    assert batch.shape == (B, 3, 64, 64)
    for i in range(B):
        one = ed.gen_img(codes[i:i + 1], labels[i:i + 1, None].numpy(),
                         noise=synth.flatten_noise([p[i:i + 1] for p in noise])).cpu()
        # same kernels, same inputs; the split-K factor of the low-resolution convs follows the batch size, so the two
        # schedules agree to the fp16 storage rounding of the activations, not bitwise
        assert float((one - batch[i]).abs().max()) < 1e-3, i
    eff = codes.clone()
    eff[1, 13], eff[2, 0] = median[13], median[0]
    ref = so.generator_forward(synthetic_sd, labels, eff, noise)
    assert float((batch - ref).norm() / ref.norm()) < 1e-3
