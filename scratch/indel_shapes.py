"""Replica of indel_tc_prepare's tile-shape choice (indel.cu) for inspection."""
import sys
def pitch8(c): return c if (c >> 3) & 1 else c + 8
def layout(NC8, tail, KCl, KC5, KCo, CinP, RA, RS, stride, SG):
    C = 8 * NC8; o = KCl * NC8 * 512 + KC5 * 2 * NC8 * 512 + NC8 * NC8 * 512 + (2 * KCo * NC8 * 512 if tail else 0) + 6 * C * 4 + (KCl + KC5) * 8
    o = (o + 15) & ~15
    PinP, PA, PC = pitch8(CinP), pitch8(C), pitch8(C)
    xs = RS * stride * PinP; as_ = (RA + 8) * PC
    w = o
    o += 2 * SG * xs * 2 + SG * RA * PA * 4 + 2 * SG * as_ * 2
    return o, w
def shapes(C=8, ks=7, down=(1, 4, 5, 5, 5, 2), L=8000, budgets=(110 * 1024, 227 * 1024)):
    ch = [C * (i + 1) for i in range(6)]; ln = []; x = L
    for s in down: x = (x - 1) // s + 1; ln.append(x)
    for step in range(11):
        dec = step >= 6; lvl = 4 - (step - 6) if dec else step
        Cin = ch[lvl + 1] if dec else (ch[lvl - 1] if lvl else 4); CinP = (Cin + 7) & ~7
        stride = 1 if dec else down[lvl]; NC8 = ch[lvl] // 8; MT = 2 if NC8 <= 2 else 1; NW = 8
        KCl = (ks * CinP + 15) // 16; KC5 = (5 * ch[lvl] + 15) // 16; KCo = (C // 8 + 1) // 2; Lout = ln[lvl]
        def shape(ra_max, sg):
            nt = (Lout + ra_max - 5) // (ra_max - 4); TP = (Lout + nt - 1) // nt
            RA = (TP + 4 + 16 * MT - 1) // (16 * MT) * 16 * MT; rows_in = (RA - 1) * stride + ks + 1; RS = (rows_in + stride - 1) // stride
            SG = sg if nt == 1 else 1
            sm, w = layout(NC8, step == 10, KCl, KC5, KCo, CinP, RA, RS, stride, SG)
            return dict(n_tiles=nt, TP=TP, RA=RA, SG=SG, smem=sm, weights=w)
        best=-1; done=None
        ra = 256
        while ra >= 16 * MT:
            sg = 1
            while sg <= 16:
                T = shape(ra, sg)
                if T["SG"] != sg or T["smem"] > 227 * 1024: break
                tiles = T["SG"] * (T["RA"] // 16); per = NW * MT
                util = tiles / (((tiles + per - 1) // per) * per); useful = Lout / (T["n_tiles"] * T["RA"])
                ctas = (227 * 1024) // T["smem"]; occ = 1.0 if ctas >= 3 else 0.9 if ctas == 2 else 0.6
                sc = useful * util * occ
                if sc > best + 1e-9: best = sc; done = dict(T, score=round(sc, 2))
                sg *= 2
            ra //= 2
        print("step %2d lvl %d NC8=%d Lout=%5d stride=%d Cin=%2d: %s" % (step, lvl, NC8, Lout, stride, Cin, done))
shapes()
