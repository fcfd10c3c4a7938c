import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from mural_b200 import PackedGenome, SiteBatch, _lib, pack_meta
L = _lib.lib()
chroms = [bench.synth_chromosome(ci) for ci in range(2)]
genome = PackedGenome({"chr%d" % (i + 1): c.tobytes() for i, c in enumerate(chroms)})
cfg, state, n_cat = bench.load_weights()
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
model = bench.build_model(cfg, state, n_cat, mode)
rng = np.random.default_rng(2718)
per = 131072
pos_l, meta_l = [], []
for ci, c in enumerate(chroms):
    span = c[: per * 220]
    idx = np.flatnonzero((span == ord("A")) | (span == ord("T")))
    sel = np.sort(rng.choice(idx, size=per, replace=False))
    pos_l.append(sel.astype(np.int32)); meta_l.append(pack_meta((span[sel] == ord("T")).astype(np.int64), np.zeros(per, np.int64), np.full(per, ci)))
sb = SiteBatch(torch.from_numpy(np.concatenate(pos_l)).cuda(), torch.from_numpy(np.concatenate(meta_l)).cuda(), genome)
with torch.no_grad():
    model.forward(None, sb); torch.cuda.synchronize()
    L.mural_profile_begin()
    for _ in range(3): model.forward(None, sb)
    torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16); L.mural_profile_end(buf, len(buf))
prof = json.loads(buf.value.decode())
tot = sum(v["ms"] for v in prof.values())
print("sparse %s: %d sites, %.3f ms per call -> %.1f M sites/s" % (mode, len(sb), tot / 3, len(sb) / (tot / 3) / 1e3))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:10]:
    print("   %-40s n=%3d %8.3f ms" % (k[:40], v["count"] // 3, v["ms"] / 3))
