"""Analytic work model of the generator forward (1 MAC = 2 FLOP), used by bench.py for the roofline figure.

dense    = what the reference executes (512-channel style map through conv_gamma/conv_beta)
factored = the exactly equivalent region-factored form this build runs (unpadded: 19 label channels, K = 171)
"""
BLOCKS = [("head_0", 16, 16, 1, True), ("G_middle_0", 16, 16, 2, True), ("G_middle_1", 16, 16, 2, True),
          ("up_0", 16, 8, 4, True), ("up_1", 8, 4, 8, True), ("up_2", 4, 2, 16, True), ("up_3", 2, 1, 32, False)]


def generator_macs(crop=256, ngf=64, label_nc=19, style_len=512, nhidden=128):
    sw = crop // 32
    dense = factored = 0
    px = sw * sw
    fc = px * 9 * label_nc * 16 * ngf
    dense += fc
    factored += fc
    for name, fi, fo, mul, styled in BLOCKS:
        fin, fout = fi * ngf, fo * ngf
        fmid = min(fin, fout)
        px = (sw * mul) ** 2
        aces = [fin, fmid] + ([fin] if fin != fout else [])
        for C in aces:
            spade = px * 9 * (label_nc * nhidden + nhidden * 2 * C)
            dense += spade
            factored += spade
            if styled:
                dense += px * 9 * style_len * 2 * C + label_nc * style_len * style_len
                factored += px * 9 * label_nc * 2 * C          # one-hot x Weff
                factored += 9 * label_nc * style_len * 2 * C    # Weff table build
                factored += label_nc * style_len * style_len    # fc_mu
        convs = px * (9 * fin * fmid + 9 * fmid * fout + (fin * fout if fin != fout else 0))
        dense += convs
        factored += convs
    img = crop * crop * 9 * ngf * 3
    return dense + img, factored + img


if __name__ == "__main__":
    for c in (256, 512):
        d, f = generator_macs(c)
        print(c, "dense %.3f GMAC = %.2f GFLOP; factored %.3f GMAC = %.2f GFLOP" % (d / 1e9, 2 * d / 1e9, f / 1e9, 2 * f / 1e9))
