"""CPU study (numpy, fp64) of preconditioners for the Poisson-blending solve — evidence for DESIGN §8 item 4, not part of
the product or of pytest.   python tests/_poisson_precond_study.py

Iterations of (preconditioned) CG to a 1e-11 relative residual on the 256x256 benchmark case (synth.make_blend_case,
67 % of the pixels unknown):
    plain CG (what csrc/blend.cu runs)                                    653
    incomplete-Poisson  M^-1 = K K^T, K = I - L D^-1 (Ament et al.)       346
    the same + additive coarse correction on 16x16 aggregates (256 dof)   110
    the same + additive coarse correction on  8x8  aggregates (1024 dof)   75
With the 256 x 256 coarse operator of the 16x16 variant inverted once per system (dense, np.linalg.inv; condition number
125) the count stays 110 whether the inverse is kept in fp64 or rounded to fp32 (256 KB per system = 32 KB per CTA of
an 8-CTA cluster), and the uint8 result still matches spsolve (1 byte of 65 536 differs, max |f - f_spsolve| 1.7e-9).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ctrlhair_b200 import synth  # noqa: E402
from oracle import blend_oracle as bo  # noqa: E402

H = W = 256


def lap(p):
    q = 4 * p
    q[1:] -= p[:-1]; q[:-1] -= p[1:]; q[:, 1:] -= p[:, :-1]; q[:, :-1] -= p[:, 1:]
    return q


def shift(r, dy, dx):
    o = np.zeros_like(r)
    ys, yd = slice(max(dy, 0), H + min(dy, 0)), slice(max(-dy, 0), H + min(-dy, 0))
    xs, xd = slice(max(dx, 0), W + min(dx, 0)), slice(max(-dx, 0), W + min(-dx, 0))
    o[yd, xd] = r[ys, xs]
    return o


def main():
    face, gen, fp, tp = synth.make_blend_case(H, W, 900)
    mask = 1 - bo.blend_mask(tp, fp)
    s, t = np.power(face.astype(float), 1 / 2.2)[:, :, 0], np.power(gen.astype(float), 1 / 2.2)[:, :, 0]
    U, m = bo.unknown_set(mask), mask != 0
    known = np.where(U, 0.0, t)
    rhs = np.where(U, np.where(m, lap(s), t) - (lap(known) - 4 * known), 0.0)

    def A(p):
        return np.where(U, lap(p), 0.0)

    def ip(r):   # K K^T r: two one-sided stencils (r is zero outside U)
        y = np.where(U, r + 0.25 * (shift(r, 1, 0) + shift(r, 0, 1)), 0.0)
        return np.where(U, y + 0.25 * (shift(y, -1, 0) + shift(y, 0, -1)), 0.0)

    def coarse(r, f):   # Galerkin coarse operator for piecewise-constant aggregates, solved by 30 inner CG steps
        Hc, Wc = H // f, W // f

        def Ac(pc):
            return A(np.where(U, np.repeat(np.repeat(pc, f, 0), f, 1), 0.0)).reshape(Hc, f, Wc, f).sum((1, 3))
        rc = r.reshape(Hc, f, Wc, f).sum((1, 3))
        e, rr, = np.zeros_like(rc), rc.copy()
        pp, g = rr.copy(), (rr * rr).sum()
        for _ in range(30):
            if g < 1e-30:
                break
            q = Ac(pp); a = g / (pp * q).sum(); e += a * pp; rr -= a * q
            gn = (rr * rr).sum(); pp = rr + (gn / g) * pp; g = gn
        return np.where(U, np.repeat(np.repeat(e, f, 0), f, 1), 0.0)

    def pcg(prec, name):
        x = np.where(U, np.where(m, s, t), 0.0)
        r = rhs - A(x); z = prec(r); p = z.copy(); rz = (r * z).sum(); bb = (rhs * rhs).sum(); it = 0
        while (r * r).sum() > 1e-22 * bb and it < 5000:
            q = A(p); a = rz / (p * q).sum(); x += a * p; r -= a * q
            z = prec(r); rzn = (r * z).sum(); p = z + (rzn / rz) * p; rz = rzn; it += 1
        print("%-70s %d iterations" % (name, it))

    pcg(lambda r: r, "plain CG")
    pcg(ip, "incomplete-Poisson K K^T")
    pcg(lambda r: ip(r) + coarse(r, 16), "incomplete-Poisson + coarse correction, 16x16 aggregates")
    pcg(lambda r: ip(r) + coarse(r, 8), "incomplete-Poisson + coarse correction, 8x8 aggregates")


if __name__ == "__main__":
    main()
