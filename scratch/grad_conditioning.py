"""Is the one-step gradient of the training test well conditioned?  fp32 vs fp64 autograd of the ORACLE itself (CPU)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLD, load_snv_golden
from oracle import network_t as NT
from oracle import encode_np as E
tag = sys.argv[1]
k = np.load(os.path.join(GOLD, "encode_kat.npz"))
genome = {str(n): str(s) for n, s in zip(k["genome_names"], k["genome_seqs"])}
z, cfg, state = load_snv_golden(tag)
n = 48
names = list(genome)
ch, st, sd = z["chrom"][:n], z["start"][:n], z["strand"][:n]
cat = np.empty((n, int(z["n_cat"])), np.int64); oh = np.empty((n, 4, 2 * cfg["distal_radius"] + 1), np.float32)
for c in range(len(names)):
    m = ch == c
    if m.any():
        sym = E.seq_to_symbols(genome[names[c]])
        cat[m] = E.kmer_windows(sym, st[m], sd[m], cfg["local_radius"], cfg["local_order"]); oh[m] = E.onehot_windows(sym, st[m], sd[m], cfg["distal_radius"])
labels = (z["start"][:n] % 4).astype(np.int64)
grads = {}
for dt in (torch.float64, torch.float32):
    sd_ = {kk: torch.tensor(np.asarray(v), dtype=dt, requires_grad=("running" not in kk)) for kk, v in state.items() if "num_batches" not in kk}
    torch.manual_seed(0)
    ref = NT.network2_forward(sd_, cat, oh, dt, train=True)
    NT.ce_sum(ref, labels).backward()
    grads[dt] = {kk: v.grad.double().numpy().reshape(-1) for kk, v in sd_.items() if v.grad is not None}
rows = []
for kk, g64 in grads[torch.float64].items():
    g32 = grads[torch.float32][kk]
    rows.append((np.abs(g32 - g64).max() / max(1e-3, np.abs(g64).max()), kk))
rows.sort(reverse=True)
for e, nm in rows[:6]:
    print("%.3e %s" % (e, nm))
