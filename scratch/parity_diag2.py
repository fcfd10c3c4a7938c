import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from conftest import load_snv_golden
from test_gpu_snv_forward import build_model
from mural_b200 import PackedGenome, SiteBatch, pack_meta
from oracle import encode_np as E, network_t as NT
chrom = bench.synth_chromosome(0)
genome = PackedGenome({"chr1": chrom.tobytes()})
st = np.array([790584, 790595, 790605, 790000], dtype=np.int32); sd = np.array([1, 1, 1, 1])
sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), genome)
z, cfg, state = load_snv_golden("ex_ckpt6")
sym = E.seq_to_symbols(chrom.tobytes().decode())
cat = E.kmer_windows(sym, st, sd, cfg["local_radius"], cfg["local_order"]); oh = E.onehot_windows(sym, st, sd, cfg["distal_radius"])
taps = {}
with torch.no_grad():
    ref = NT.network2_forward(state, cat, oh, torch.float32, taps=taps)
for mode in ("fp32", "bf16"):
    m = build_model(cfg, state, int(z["n_cat"]), mode=mode)
    m.set_debug(True, chunk=0)
    with torch.no_grad():
        out = m.forward(None, sb)
    print(mode, "probs", torch.softmax(out, 1).cpu().numpy().round(4).tolist()[0], "ref", torch.softmax(ref, 1).numpy().round(4).tolist()[0])
    for name in ("pool1", "pool1_2", "rb1", "rb1_2", "rb2", "rb2_2", "gmax", "gmax_2", "logit_mid", "logit_large", "logit_local"):
        if name not in taps: continue
        r = np.asarray(taps[name])
        g = m.debug_tap(name)
        if r.ndim == 3: g = g.reshape(len(st), -1, r.shape[1]).transpose(0, 2, 1)
        else: g = g.reshape(r.shape)
        for i in (0, 3):
            print("   %-12s site %d: max|ref| %.3g  max|err| %.3g  rel %.3g" % (name, i, np.abs(r[i]).max(), np.abs(g[i] - r[i]).max(), np.abs(g[i] - r[i]).max() / max(1e-9, np.abs(r[i]).max())))
# BN conditioning of the checkpoint
for k in sorted(state):
    if k.endswith("running_var") and ("RBs" in k or "conv" in k) and ".layer." not in k:
        v = state[k]; g = state[k.replace("running_var", "weight")]; mu = state[k.replace("running_var", "running_mean")]
        a = g / np.sqrt(v + 1e-5)
        print("%-28s |a| max %.3g  |mu|/sigma max %.3g" % (k[:-12], np.abs(a).max(), (np.abs(mu) / np.sqrt(v + 1e-5)).max()))
