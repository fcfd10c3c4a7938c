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
rng = np.random.default_rng(7)
n = 400000
st = np.sort(rng.choice(np.arange(15000, 2_000_000), n, replace=False)).astype(np.int32)
sd = rng.integers(0, 2, n)
sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), genome)
z, cfg, state = load_snv_golden("ex_ckpt6")
m = build_model(cfg, state, int(z["n_cat"]), mode="fp32")
res = {}
with torch.no_grad():
    res["fp32"] = torch.softmax(m.forward(None, sb), 1)
    m.compute_mode = "bf16"
    for key, envs in (("bf16", {}), ("nolat", {"MURAL_NO_LATTICE": "1"}), ("nodense", {"MURAL_NO_DENSE_STEM": "1"}), ("notail", {"MURAL_NO_TAIL": "1"}),
                      ("v2", {"MURAL_TC_V2": "1"}), ("nomlp", {"MURAL_NO_MLP_TC": "1"})):
        for k, v in envs.items(): os.environ[k] = v
        res[key] = torch.softmax(m.forward(None, sb), 1)
        for k in envs: os.environ.pop(k)
d = (res["fp32"] - res["bf16"]).abs().max(1).values
bad = (d > 5e-3).nonzero().flatten().cpu().numpy()
print("bad sites:", bad.tolist(), st[bad].tolist(), sd[bad].tolist())
for key in res:
    print(key, [(res[key][i].cpu().numpy().round(4).tolist()) for i in bad[:3]])
sym = E.seq_to_symbols(chrom.tobytes().decode())
cat = E.kmer_windows(sym, st[bad], sd[bad], cfg["local_radius"], cfg["local_order"]); oh = E.onehot_windows(sym, st[bad], sd[bad], cfg["distal_radius"])
with torch.no_grad():
    print("oracle", torch.softmax(NT.network2_forward(state, cat, oh, torch.float32), 1).numpy().round(4).tolist())
for i in bad[:3]:
    w = chrom[st[i] - 210: st[i] + 211].tobytes().decode()
    print(i, st[i], sd[i], "N in window:", w.count("N"), "window N span:", (w.find("N"), w.rfind("N")))
