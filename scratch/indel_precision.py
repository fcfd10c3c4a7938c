"""CPU emulation: UNet_Small with BatchNorm folded into the conv weights and conv operands rounded to a short format
(what a tensor-core conv would see).  Reports max |out - ref| / max(1, scale) per golden for each variant."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.nn.functional as F
from oracle import encode_np as E
GOLD = os.path.join(ROOT, "tests", "golden")
k = np.load(os.path.join(GOLD, "encode_kat.npz"))
genome = {str(n): str(s) for n, s in zip(k["genome_names"], k["genome_seqs"])}
names = list(genome); syms = [E.seq_to_symbols(genome[n]) for n in names]

def rnd(x, mode):
    if mode == "f32": return x
    if mode == "bf16": return x.to(torch.bfloat16).to(torch.float32)
    if mode == "f16": return x.to(torch.float16).to(torch.float32)
    if mode == "bf16x2":
        h = x.to(torch.bfloat16).to(torch.float32); return h + (x - h).to(torch.bfloat16).to(torch.float32)
    if mode == "f16x2":
        h = x.to(torch.float16).to(torch.float32); return h + (x - h).to(torch.float16).to(torch.float32)
    raise ValueError(mode)

def fold(sd, conv, bn, has_bias=True):
    W = torch.from_numpy(np.asarray(sd[conv + ".weight"])).double()
    b = torch.from_numpy(np.asarray(sd[conv + ".bias"])).double() if has_bias and (conv + ".bias") in sd else torch.zeros(W.shape[0], dtype=torch.float64)
    if bn:
        g = torch.from_numpy(np.asarray(sd[bn + ".weight"])).double(); be = torch.from_numpy(np.asarray(sd[bn + ".bias"])).double()
        m = torch.from_numpy(np.asarray(sd[bn + ".running_mean"])).double(); v = torch.from_numpy(np.asarray(sd[bn + ".running_var"])).double()
        a = g / torch.sqrt(v + 1e-5)
        W = W * a.view(-1, 1, 1); b = a * b + be - m * a
    return W.float(), b.float()

def run(sd, oh, down, use_reverse, xa, wa):
    def conv(x, W, b, stride=1):
        return F.conv1d(rnd(x, xa).double(), rnd(W, wa).double(), None, stride=stride, padding=(W.shape[2] - 1) // 2).float() + b.view(1, -1, 1)
    o = torch.from_numpy(oh)
    if use_reverse:
        W, b = fold(sd, "conv.0", "conv.1")
        f = lambda z: F.conv1d(z, W, b, padding=(W.shape[2] - 1) // 2)      # table lookups: exact fp32
        o = f(o) + f(o.flip([1, 2])).flip([2])
    def cblock(x, p):
        W1, b1 = fold(sd, p + ".conv.0", p + ".conv.1", False); W2, b2 = fold(sd, p + ".conv.3", p + ".conv.4", False)
        return x + conv(F.silu(conv(x, W1, b1)), W2, b2)
    enc = []
    for i in range(6):
        W, b = fold(sd, "uplblocks.%d.0" % i, "uplblocks.%d.1" % i)
        o = cblock(conv(o, W, b, down[i]), "upblocks.%d.0" % i); enc.append(o)
    for i in range(5):
        o = F.interpolate(o, scale_factor=float(down[5 - i]), mode="nearest")
        W, b = fold(sd, "downlblocks.%d.1" % i, "downlblocks.%d.2" % i)
        o = enc[4 - i] + cblock(conv(o, W, b), "downblocks.%d.0" % i)
    W, b = fold(sd, "out_conv.0", "out_conv.1"); o = F.relu(conv(o, W, b))
    W, b = fold(sd, "out_conv.3", None); o = F.softplus(conv(o, W, b))
    o = o.max(dim=2)[0]
    g, be, m, v = [torch.from_numpy(np.asarray(sd["out_fc.0." + s])) for s in ("weight", "bias", "running_mean", "running_var")]
    o = (o - m) / torch.sqrt(v + 1e-5) * g + be
    return F.softplus(F.linear(o, torch.from_numpy(np.asarray(sd["out_fc.2.weight"])), torch.from_numpy(np.asarray(sd["out_fc.2.bias"]))))

tags = sys.argv[1:] or ["hs_ins", "hs_del_start", "ex_indel9", "at_ins", "dm_del_end", "mm_ins"]
variants = [("f32", "f32"), ("bf16", "bf16"), ("f16", "f16"), ("f16x2", "f16"), ("f16", "f16x2"), ("bf16x2", "bf16x2"), ("f16x2", "f16x2")]
for tag in tags:
    z = np.load(os.path.join(GOLD, "indel_%s.npz" % tag))
    sd = {kk[2:]: z[kk] for kk in z.files if kk.startswith("w:")}
    ch, st, sdn = z["chrom"], z["start"], z["strand"]; Rd = int(z["distal_radius"])
    n = min(len(st), 48)
    oh = np.empty((n, 4, 2 * Rd), np.float32)
    for c in range(len(names)):
        m = ch[:n] == c
        oh[m] = E.onehot_windows(syms[c], st[:n][m], sdn[:n][m], Rd, "indel")
    ref = z["ref_out"][:n]
    res = []
    with torch.no_grad():
        for xa, wa in variants:
            o = run(sd, oh, [int(v) for v in z["down"]], bool(z["use_reverse"]), xa, wa).numpy()
            res.append("x=%s,w=%s: %.2e" % (xa, wa, np.abs(o - ref).max() / max(1.0, np.abs(ref).max())))
    print(tag, "Rd=%d scale=%.2f" % (Rd, np.abs(ref).max()), " | ".join(res), flush=True)
