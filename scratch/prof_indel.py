import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import mural_b200._lib as _L0
if os.environ.get("MURAL_LIB"): _L0.LIB_PATH = os.environ["MURAL_LIB"]
from mural_b200 import PackedGenome, SiteBatch, _lib, model_choice, pack_meta
L = _lib.lib()
chroms = [bench.synth_chromosome(0)]
genome = PackedGenome({"chr1": chroms[0].tobytes()})
z = np.load(os.path.join(ROOT, "tests", "golden", "indel_hs_ins.npz"))
state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
cfg = {"CNN_out_channels": state["uplblocks.0.0.weight"].shape[0], "CNN_kernel_size": state["uplblocks.0.0.weight"].shape[2],
       "down_list": [int(v) for v in z["down"]], "use_reverse": bool(z["use_reverse"]), "n_class": state["out_fc.2.weight"].shape[0]}
m = model_choice(0, cfg, {"n_class": cfg["n_class"]}, "indel")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
m.to("cuda").eval()
n = int(os.environ.get("NSITES", "2048"))
pos = torch.from_numpy((20000 + 50 * np.arange(n)).astype(np.int32)).cuda()
meta = torch.from_numpy(pack_meta(np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64))).cuda()
sb = SiteBatch(pos, meta, genome)
with torch.no_grad():
    for _ in range(2): m.forward(sb, distal_radius=4000)
    torch.cuda.synchronize()
    L.mural_profile_begin()
    m.forward(sb, distal_radius=4000)
    torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
L.mural_profile_end(buf, len(buf))
prof = json.loads(buf.value.decode())
tot = sum(v["ms"] for v in prof.values()); nl = sum(v["count"] for v in prof.values())
print("%d launches, %.3f ms per %d-site batch -> %.0f sites/s" % (nl, tot, n, n / tot * 1e3))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]:
    print("   %-44s n=%4d  %8.3f ms" % (k[:44], v["count"], v["ms"]))

# ---- training step (batch 32)
from mural_b200.training import IndelTrainState
m.train()
ts = IndelTrainState(m, 4000, "Adam", lr=1e-4, weight_decay=1e-5, seed=0)
B = 32
rng = np.random.default_rng(1)
lab = rng.choice(8, size=B * 4, p=[0.907] + [0.093 / 7] * 7)
pos = torch.from_numpy((20000 + 50 * np.arange(B * 4)).astype(np.int32)).cuda()
meta = torch.from_numpy(pack_meta(np.zeros(B * 4, np.int64), lab, np.zeros(B * 4, np.int64))).cuda()
for i in range(2):
    ts.step(SiteBatch(pos[i * B:(i + 1) * B], meta[i * B:(i + 1) * B], genome))
torch.cuda.synchronize()
L.mural_profile_begin()
for i in range(2, 4):
    ts.step(SiteBatch(pos[i * B:(i + 1) * B], meta[i * B:(i + 1) * B], genome))
torch.cuda.synchronize()
buf = C.create_string_buffer(1 << 16)
L.mural_profile_end(buf, len(buf))
prof = json.loads(buf.value.decode())
tot = sum(v["ms"] for v in prof.values()); nl = sum(v["count"] for v in prof.values())
print("TRAIN: %d launches / 2 steps, %.3f ms kernel time per batch-32 step" % (nl, tot / 2))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]:
    print("   %-44s n=%4d  %8.3f ms/step" % (k[:44], v["count"] // 2, v["ms"] / 2))
