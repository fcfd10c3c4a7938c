"""CPU experiment (no GPU): which bf16 rounding of the tcgen05 path dominates its error?

Emulates the data flow of snv_tc.cu / snv_tail.cu in float64 with optional bf16 rounding of
  W   the BN-folded conv weights (a * W),
  A   the activation operand of every conv inside a stage (relu(x) in bf16),
  IO  what is stored in bf16 between kernels: stem rows x0, stage outputs z1 / z2, the parked conv2 jump,
and compares the probabilities with the unrounded float64 network on (i) random sites and (ii) sites whose window touches an
N run (where `ex_ckpt6` showed 0.19 outliers).  Uses oracle/ (this is a scratch diagnostic, not product code).

    python scratch/bf16_error_sources.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.nn.functional as F

import bench
from conftest import load_snv_golden
from oracle import encode_np as E

D = torch.float64
EPS = 1e-5


def rb(x, on):
    return x.to(torch.bfloat16).to(D) if on else x


def T(sd, k):
    return torch.from_numpy(np.asarray(sd[k])).to(D)


def bn_ab(sd, p):
    a = T(sd, p + ".weight") / torch.sqrt(T(sd, p + ".running_var") + EPS)
    return a, T(sd, p + ".bias") - T(sd, p + ".running_mean") * a


def conv_folded(act, sd, bn, cv, fl):
    """conv(BN(act)) the way the kernels do it: (a*W in bf16) x (act in bf16) + exact constant part (bias, edge-corrected)."""
    a, b = bn_ab(sd, bn)
    W, bias = T(sd, cv + ".weight"), T(sd, cv + ".bias")
    Wf = rb(W * a.view(1, -1, 1), fl["W"])
    const = F.conv1d(b.view(1, -1, 1).expand(1, -1, act.shape[2]).contiguous(), W, bias, padding=1)
    return F.conv1d(rb(act, fl["A"]), Wf, None, padding=1) + const


def resblocks(R, sd, prefix, fl):
    for i in range(2):
        p = "%s.%d" % (prefix, i)
        t = conv_folded(F.relu(R), sd, p + ".bn1", p + ".conv1", fl)
        R = R + conv_folded(F.relu(t), sd, p + ".bn2", p + ".conv2", fl)
    return R


def branch(x, sd, sfx, pools, fl):
    a, b = bn_ab(sd, "conv1%s.0" % sfx)
    o = F.conv1d(x * a.view(1, -1, 1) + b.view(1, -1, 1), T(sd, "conv1%s.1.weight" % sfx), T(sd, "conv1%s.1.bias" % sfx), padding=1)
    x0 = rb(F.max_pool1d(o, *pools[0]), fl["IO"])
    z1 = rb(resblocks(x0, sd, "RBs1" + sfx, fl) + x0, fl["IO"])
    j = conv_folded(F.max_pool1d(z1, *pools[1]), sd, "conv2%s.0" % sfx, "conv2%s.1" % sfx, fl)
    z2 = rb(resblocks(j, sd, "RBs2" + sfx, fl) + rb(j, fl["IO"]), fl["IO"])
    h = F.relu(conv_folded(F.max_pool1d(z2, *pools[2]), sd, "conv3%s.0" % sfx, "conv3%s.1" % sfx, fl)).max(2)[0]
    fc = "distal_fc1" if sfx == "" else "distal_fc2"
    a, b = bn_ab(sd, fc + ".0")
    return F.linear(h * a + b, T(sd, fc + ".2.weight"), T(sd, fc + ".2.bias"))


def forward(sd, cat, oh, fl):
    lo = T(sd, "emb_layer.weight")[torch.from_numpy(cat)].reshape(len(cat), -1)
    i = 0
    while ("lin_layers.%d.weight" % i) in sd:
        lo = F.relu(F.linear(lo, T(sd, "lin_layers.%d.weight" % i), T(sd, "lin_layers.%d.bias" % i)))
        a, b = bn_ab(sd, "bn_layers.%d" % i)
        lo = lo * a + b
        i += 1
    lo = F.linear(lo, T(sd, "local_fc.0.weight"), T(sd, "local_fc.0.bias"))
    x = torch.from_numpy(oh).to(D)
    L = x.shape[2]
    d1 = branch(x[:, :, L // 2 - 100: L // 2 + 101], sd, "", ((3, 3, 1), (3, 3, 1), (3, 3, 1)), fl)
    d2 = branch(x, sd, "_2", ((15, 15, 7), (7, 7, 3), (3, 3, 1)), fl)
    return (F.softmax(lo, 1) + (F.softmax(d1, 1) + F.softmax(d2, 1)) / 2) / 2


def main():
    chrom = bench.synth_chromosome(0)
    sym = E.seq_to_symbols(chrom.tobytes().decode())
    isN = chrom == ord("N")
    edges = np.flatnonzero(np.diff(isN.astype(np.int8)) != 0)
    edges = edges[(edges > 30000) & (edges < len(chrom) - 30000)]
    rng = np.random.default_rng(3)
    for tag in ("ex_ckpt6", "hs_AT"):
        z, cfg, sd = load_snv_golden(tag)
        R = cfg["distal_radius"]
        near = np.unique(np.concatenate([e + rng.integers(-R, R, 40) for e in edges[:16]]))
        near = near[~isN[near]][:400]
        rand = rng.integers(50000, 2_000_000, 3000)
        rand = rand[~isN[rand]][:1200 if R <= 200 else 300]
        print("== %s (R_d=%d)" % (tag, R))
        for name, pos in (("windows touching an N run", near), ("random sites", rand)):
            pos = np.sort(pos).astype(np.int64)
            sd_ = rng.integers(0, 2, len(pos))
            cat = E.kmer_windows(sym, pos, sd_, cfg["local_radius"], cfg["local_order"])
            oh = E.onehot_windows(sym, pos, sd_, R)
            with torch.no_grad():
                ref = forward(sd, cat, oh, {"W": 0, "A": 0, "IO": 0})
                row = []
                for fl in ({"W": 1, "A": 0, "IO": 0}, {"W": 0, "A": 1, "IO": 0}, {"W": 0, "A": 0, "IO": 1}, {"W": 1, "A": 1, "IO": 1}):
                    d = (forward(sd, cat, oh, fl) - ref).abs().max(1)[0]
                    row.append((float(d.max()), float(d.median())))
            print("   %-28s n=%4d  max|dp| (median):  W only %.2e (%.1e) | A only %.2e (%.1e) | IO only %.2e (%.1e) | all %.2e (%.1e)" %
                  ((name, len(pos)) + tuple(v for r in row for v in r)))


if __name__ == "__main__":
    main()
