"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy / scipy) of the reference's post-processing either side of the
generator.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import it.

  poisson_blending          poisson_blending.py:15-87      (sparse system assembled vectorised, solved with spsolve)
  blend_mask                hair_editor.py:297-306         (hair-region union, 13x13 / 5x5 elliptical dilation)
  postprocess_blending      hair_editor.py:257-308
  rgb_to_hsv_u8 / hsv_to_rgb_u8   ui/backend.py:98-101,108-125  (cv2.cvtColor on uint8: OpenCV's 8-bit algorithms,
                            third-party dependency absent from /root/reference: opencv-python, 4.13.0 in this image;
                            restated from its published fixed-point / float formulas and pinned exhaustively against
                            cv2 by oracle/make_golden_blend.py -> tests/golden/blend.npz digests)
  mask_one_hot_to_label / mask_label_to_one_hot / split_hair_face   shape_branch/shape_util.py:6-26

Parity pin: oracle/make_golden_blend.py imports the unmodified reference `poisson_blending` and cv2 and stores their
outputs on seeded inputs in tests/golden/blend.npz; tests/test_blend.py checks this file against them.
"""
import hashlib

import numpy as np
import scipy.sparse
from scipy.sparse.linalg import spsolve

HAIR_IDX = 13          # global_value_utils.py:49-52
BACKGROUND_IDX = 0     # PARSING_LABEL_LIST.index('background')

# cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k)): half-widths of each row around the centre column
# (rows are symmetric about the centre column; checked against cv2 by make_golden_blend.py)
ELLIPSE_HALF_WIDTH = {13: (0, 3, 4, 5, 6, 6, 6, 6, 6, 5, 4, 3, 0), 5: (0, 2, 2, 2, 0)}


def ellipse_kernel(k):
    hw = ELLIPSE_HALF_WIDTH[k]
    se = np.zeros((k, k), np.uint8)
    for i, w in enumerate(hw):
        se[i, k // 2 - w:k // 2 + w + 1] = 1
    return se


def dilate(mask, k):
    """cv2.dilate(mask, ellipse k x k, iterations=1): anchor at the centre, pixels outside the image never win
    (the default border of the morphology filters), so the result is the OR over the in-image footprint."""
    se = ellipse_kernel(k)
    H, W = mask.shape
    r = k // 2
    pad = np.zeros((H + 2 * r, W + 2 * r), np.uint8)
    pad[r:r + H, r:r + W] = mask
    out = np.zeros((H, W), np.uint8)
    for dy in range(k):
        for dx in range(k):
            if se[dy, dx]:
                out = np.maximum(out, pad[dy:dy + H, dx:dx + W])
    return out


def blend_mask(target_parsing, face_parsing):
    """hair_editor.py:297-306 -> res_mask_dilated uint8 [H, W] (1 = take the generated image as it is)."""
    target_parsing = np.asarray(target_parsing)
    face_parsing = np.asarray(face_parsing)
    res_mask = np.logical_or(target_parsing == HAIR_IDX, face_parsing == HAIR_IDX).astype(np.uint8)
    d13 = dilate(res_mask, 13)
    d5 = dilate(res_mask, 5)
    bg = (target_parsing == BACKGROUND_IDX).astype(np.uint8)
    return (d13 * (1 - bg) + d5 * bg).astype(np.uint8)


def laplacian_rows(H, W):
    """poisson_blending.py:15-27: 4 on the diagonal, -1 for each 4-neighbour inside the image (rows are truncated at
    the image border, the blocks do not wrap)."""
    n = H * W
    idx = np.arange(n).reshape(H, W)
    rows = [idx.ravel()]
    cols = [idx.ravel()]
    vals = [np.full(n, 4.0)]
    for a, b in ((idx[:, :-1], idx[:, 1:]), (idx[:, 1:], idx[:, :-1]), (idx[:-1, :], idx[1:, :]), (idx[1:, :], idx[:-1, :])):
        rows.append(a.ravel())
        cols.append(b.ravel())
        vals.append(np.full(a.size, -1.0))
    return scipy.sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def unknown_set(mask):
    """Pixels whose row of the system stays a Laplacian row: mask != 0, plus every pixel of the image border
    (the identity-row loop of poisson_blending.py:50-58 only visits 1..H-2 x 1..W-2)."""
    m = np.asarray(mask).reshape(mask.shape[0], mask.shape[1]) != 0
    u = m.copy()
    u[0, :] = u[-1, :] = True
    u[:, 0] = u[:, -1] = True
    return u


def poisson_system(source_c, target_c, mask):
    """One channel: (A, b) exactly as poisson_blending.py:45-76 builds them."""
    H, W = source_c.shape
    lap = laplacian_rows(H, W)
    m = (np.asarray(mask).reshape(H, W) != 0)
    u = unknown_set(m).ravel()
    eye = scipy.sparse.identity(H * W, format="csr")
    d_u = scipy.sparse.diags(u.astype(np.float64))
    d_k = scipy.sparse.diags((~u).astype(np.float64))
    A = (d_u @ lap + d_k @ eye).tocsc()
    b = lap.dot(source_c.ravel())
    mf = m.ravel()
    b[~mf] = target_c.ravel()[~mf]
    return A, b


def poisson_solve(source, target, mask, with_gamma=True):
    """Gamma-domain solution (float64 [H, W, C]) before the final power / clip / uint8 truncation."""
    gamma = 2.2 if with_gamma else 1
    s = np.power(np.asarray(source).astype("float"), 1 / gamma)
    t = np.power(np.asarray(target).astype("float"), 1 / gamma)
    res = t.copy()
    for c in range(s.shape[2]):
        A, b = poisson_system(s[:, :, c], t[:, :, c], mask)
        res[:, :, c] = spsolve(A, b).reshape(s.shape[:2])
    return res


def finish(res, with_gamma=True):
    """poisson_blending.py:81-86."""
    gamma = 2.2 if with_gamma else 1
    with np.errstate(invalid="ignore"):
        res = np.power(res, gamma)
    res = np.where(np.isnan(res), 0.0, res)   # negative ** 2.2 = nan in numpy; `res < 0` is False for nan and
    res[res > 255] = 255                      # astype('uint8') of nan is 0 on x86
    res[res < 0] = 0
    return res.astype("uint8")


def gamma_tables():
    """v ** (1/2.2) for the 256 levels and uint8((v ** (1/2.2)) ** 2.2) as THIS host's numpy computes them.  pow()
    differs in the last bit between numpy builds (SVML vs libm); an untouched pixel v returns as v or v - 1 accordingly
    (here: 2 -> 1 and 7 -> 6 with the AVX-512 loops).  The CUDA path takes these tables from its caller."""
    fwd = np.power(np.arange(256).astype("float"), 1 / 2.2)
    return fwd, finish(fwd.copy(), True)


def poisson_blending(source, target, mask, with_gamma=True):
    """source, target uint8 [H, W, 3]; mask [H, W(, 1)], non-zero = solve for the source's gradients there."""
    return finish(poisson_solve(source, target, mask, with_gamma), with_gamma)


def tensor_to_cv2_u8(res_img):
    """hair_editor.py:273-288 for a generator output [3, H, W] in [-1, 1]: HWC, *127.5+127.5, astype(uint8)."""
    res = np.transpose(np.asarray(res_img, dtype=np.float32), [1, 2, 0])
    res = res * 127.5 + 127.5
    return res.astype("uint8")


def postprocess_blending(face_img, res_img, face_parsing, target_parsing, blending=True):
    """hair_editor.py:257-308.  face_img uint8 [H, W, 3]; res_img float [3, H, W] in [-1, 1]; parsings [H, W]."""
    res = tensor_to_cv2_u8(res_img)
    if not blending:
        return res, None
    rmd = blend_mask(target_parsing, face_parsing)[..., None]
    out = poisson_blending(np.asarray(face_img).astype("uint8"), res, 1 - rmd, with_gamma=True)
    return out, rmd


# ----------------------------------------------------------------------------------------- colour space (cv2, 8 bit)
def float_to_u8_trunc(c):
    """ndarray.astype('uint8') of a float array as x86 numpy does it: truncate toward zero, wrap modulo 256."""
    return (np.trunc(np.asarray(c, dtype=np.float64)).astype(np.int64) & 0xFF).astype(np.uint8)


def _round_half_even(x):
    return np.rint(x).astype(np.int64)


_HSV_SHIFT = 12
_SDIV = np.zeros(256, np.int64)
_HDIV180 = np.zeros(256, np.int64)
_SDIV[1:] = _round_half_even((255 << _HSV_SHIFT) / (1.0 * np.arange(1, 256)))
_HDIV180[1:] = _round_half_even((180 << _HSV_SHIFT) / (6.0 * np.arange(1, 256)))


def rgb_to_hsv_u8(rgb):
    """cv2.cvtColor(uint8 [...,3], COLOR_RGB2HSV): OpenCV's RGB2HSV_b fixed-point path, H in 0..179."""
    rgb = np.asarray(rgb, dtype=np.uint8).astype(np.int64)
    r, g, b = rgb[..., 0], rgb[..., 1], rgb[..., 2]
    v = np.maximum(np.maximum(r, g), b)
    vmin = np.minimum(np.minimum(r, g), b)
    diff = v - vmin
    vr = v == r
    vg = v == g
    s = (diff * _SDIV[v] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    h = np.where(vr, g - b, np.where(vg, b - r + 2 * diff, r - g + 4 * diff))
    h = (h * _HDIV180[diff] + (1 << (_HSV_SHIFT - 1))) >> _HSV_SHIFT
    h = h + np.where(h < 0, 180, 0)
    return np.stack([h, s, v], -1).astype(np.uint8)


def hsv_to_rgb_u8(hsv):
    """cv2.cvtColor(uint8 [1,1,3], COLOR_HSV2RGB) — the scalar path the reference's one-pixel calls take: OpenCV's
    HSV2RGB_b = float32 HSV2RGB_native on (h, s/255, v/255) with hscale 6/180, then saturate_cast<uchar>(x*255)
    (round half to even).  The shipped binary contracts `1 - s*h` and `1 - s*(1-h)` into fused multiply-adds
    (found by exhaustive comparison); they are emulated here through float64 (the product of two float32 is exact)."""
    hsv = np.asarray(hsv, dtype=np.uint8)
    f = np.float32
    h = hsv[..., 0].astype(f)
    s = hsv[..., 1].astype(f) * f(1.0 / 255.0)
    v = hsv[..., 2].astype(f) * f(1.0 / 255.0)
    h = h * f(6.0 / 180.0)
    h = np.fmod(h, f(6.0)).astype(f)
    sector = np.floor(h).astype(np.int64)
    h = (h - sector.astype(f)).astype(f)
    bad = (sector < 0) | (sector >= 6)
    sector = np.where(bad, 0, sector)
    h = np.where(bad, f(0), h).astype(f)
    one = f(1.0)
    d = np.float64

    def one_minus_prod(a, b):   # fmaf(-a, b, 1)
        return (d(1.0) - a.astype(d) * b.astype(d)).astype(f)
    tab = np.stack([v, (v * (one - s)).astype(f), (v * one_minus_prod(s, h)).astype(f),
                    (v * one_minus_prod(s, (one - h).astype(f))).astype(f)], -1)
    sector_data = np.array([[1, 3, 0], [1, 0, 2], [3, 0, 1], [0, 2, 1], [0, 1, 3], [2, 1, 0]])
    sel = sector_data[sector]                         # [..., 3] -> (b, g, r) table slots
    b = np.take_along_axis(tab, sel[..., 0:1], -1)[..., 0]
    g = np.take_along_axis(tab, sel[..., 1:2], -1)[..., 0]
    r = np.take_along_axis(tab, sel[..., 2:3], -1)[..., 0]
    gray = hsv[..., 1] == 0
    b, g, r = (np.where(gray, v, c) for c in (b, g, r))
    out = np.stack([r, g, b], -1).astype(f) * f(255.0)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def table_digest(arr):
    return hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()


def table_digest_of(fn, triples, chunk=1 << 20):
    """SHA-256 of fn(triples) evaluated in chunks (same digest as table_digest(fn(triples)), a fraction of the memory)."""
    h = hashlib.sha256()
    for i in range(0, len(triples), chunk):
        h.update(np.ascontiguousarray(fn(triples[i:i + chunk])).tobytes())
    return h.hexdigest()


def all_rgb():
    v = np.arange(256, dtype=np.uint8)
    return np.stack(np.meshgrid(v, v, v, indexing="ij"), -1).reshape(-1, 3)


# ----------------------------------------------------------------------------------------- label maps
def mask_label_to_one_hot(img, nc=19):
    """shape_util.py:6-14: uint8 [B,1,H,W] (255 = no label) -> float32 [B,19,H,W]."""
    img = np.asarray(img).astype(np.int64).copy()
    img[img == 255] = nc
    B, _, H, W = img.shape
    out = np.zeros((B, nc + 1, H, W), np.float32)
    np.put_along_axis(out, img, 1.0, axis=1)
    return out[:, :-1]


def mask_one_hot_to_label(one_hot):
    """shape_util.py:17-20: argmax over channels (first maximum wins), 255 where every channel is 0."""
    one_hot = np.asarray(one_hot)
    lab = np.argmax(one_hot, axis=1)
    lab[one_hot.max(axis=1) == 0] = 255
    return lab


def split_hair_face(mask):
    """shape_util.py:23-26."""
    return mask[:, [HAIR_IDX]], np.concatenate([mask[:, :HAIR_IDX], mask[:, HAIR_IDX + 1:]], 1)
