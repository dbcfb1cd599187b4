"""Test helper: torch-on-CPU emulation of the packed BiSeNet schedule of csrc/bisenet.cu (same tensors, same order of
operations, fp32 arithmetic) — checks ctrlhair_b200.bisenet.pack_bisenet's folds and layouts without a GPU."""
import torch
import torch.nn.functional as F


def _unpack(w, C, taps):
    k = 3 if taps == 9 else 1
    return w.float().reshape(w.shape[0], k, k, C).permute(0, 3, 1, 2).contiguous()


def emulate(packed, img_u8, round16=False):
    """img_u8 uint8 [B,S,S,3] -> logits [B,19,S/8,S/8] (before the bilinear upsample)."""
    r16 = (lambda t: t.half().float()) if round16 else (lambda t: t)
    x = torch.as_tensor(img_u8).permute(0, 3, 1, 2).float() / 255.0
    x = (x - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    sw = packed["stem.w"][:, :147].reshape(64, 7, 7, 3).permute(0, 3, 1, 2)
    x = r16(F.relu(F.conv2d(x, sw, packed["stem.b"], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)

    def conv(name, segs, relu=True, n=None):
        acc = None
        for i, (a, C, taps) in enumerate(segs):
            w = _unpack(packed["%s.w%d" % (name, i)], C, taps)
            y = F.conv2d(a, w, None, 1, 1 if taps == 9 else 0)
            acc = y if acc is None else acc + y
        acc = acc + packed[name + ".b"].view(1, -1, 1, 1)
        if n is not None:
            acc = acc[:, :n]
        return F.relu(acc) if relu else acc

    feats = {}
    chans = [64, 128, 256, 512]
    Cin = 64
    for li in range(4):
        C = chans[li]
        for bi in range(2):
            q = "layer%d.%d" % (li + 1, bi)
            if li > 0 and bi == 0:
                a = r16(conv(q + ".conv1", [(x, Cin, 9)]))[:, :, ::2, ::2]
                xs = x[:, :, ::2, ::2]
                x = r16(conv(q + ".conv2", [(a, C, 9), (xs, Cin, 1)]))
            else:
                a = r16(conv(q + ".conv1", [(x, C, 9)]))
                x = r16(conv(q + ".conv2", [(a, C, 9), (x, C, 1)]))
            Cin = C
        feats[li] = x
    feat8, feat16, feat32 = feats[1], feats[2], feats[3]

    def dense(v, w, b, act):
        y = v @ packed[w].t() + (packed[b] if b else 0)
        return F.relu(y) if act == 1 else (torch.sigmoid(y) if act == 2 else y)

    avg = dense(feat32.mean((2, 3)), "conv_avg.w", "conv_avg.b", 1)
    a32 = r16(conv("arm32.conv", [(feat32, 512, 9)]))
    att32 = dense(a32.mean((2, 3)), "arm32.att.w", "arm32.att.b", 2)
    s32 = r16(a32 * att32[:, :, None, None] + avg[:, :, None, None])
    s32 = s32.repeat_interleave(2, 2).repeat_interleave(2, 3)
    h32 = r16(conv("conv_head32", [(s32, 128, 9)]))
    a16 = r16(conv("arm16.conv", [(feat16, 256, 9)]))
    att16 = dense(a16.mean((2, 3)), "arm16.att.w", "arm16.att.b", 2)
    s16 = r16(a16 * att16[:, :, None, None] + h32).repeat_interleave(2, 2).repeat_interleave(2, 3)
    h16 = r16(conv("conv_head16", [(s16, 128, 9)]))
    ff = r16(conv("ffm.convblk", [(feat8, 128, 1), (h16, 128, 1)]))
    attf = dense(dense(ff.mean((2, 3)), "ffm.conv1.w", None, 1), "ffm.conv2.w", None, 2)
    fo = r16(ff * (attf[:, :, None, None] + 1.0))
    of = r16(conv("conv_out.conv", [(fo, 256, 9)]))
    return conv("conv_out.conv_out", [(of, 256, 1)], relu=False, n=19)
