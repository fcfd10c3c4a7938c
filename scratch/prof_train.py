import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from mural_b200 import PackedGenome, SiteBatch, _lib, model_choice, pack_meta, weights_init
from mural_b200.training import TrainState
L = _lib.lib()
chroms = [bench.synth_chromosome(0)]
genome = PackedGenome({"chr1": chroms[0].tobytes()})
pos, meta = bench.rank_sites(chroms, 0, 1, 200000)
cfg = {"local_radius": 10, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": 1000,
       "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": 3, "CNN_out_channels": 32, "distal_fc_dropout": .25, "n_class": 4, "model_no": 2}
torch.manual_seed(0)
model = model_choice(2, cfg, dict(emb_dims=[(65, 2)] * 19, n_cont=0, n_class=4, distal_order=1, in_channels=4), "snv")
model.apply(weights_init); model.to("cuda").train()
ts = TrainState(model, "Adam", lr=1e-3, weight_decay=1e-5, use_graph=False)   # eager: the event profiler hooks the launches
rng = np.random.default_rng(0)
for B in (128, 4096):
    sel = np.sort(rng.choice(len(pos), size=B * 8, replace=False))
    lab = rng.choice(4, size=len(sel), p=[0.952381, 0.0140095, 0.0198, 0.0138095])
    mt = pack_meta(meta[sel] & 1, lab, meta[sel] >> 8)
    dp, dm = torch.from_numpy(pos[sel]).cuda(), torch.from_numpy(mt).cuda()
    for i in range(3):
        ts.step(SiteBatch(dp[i * B:(i + 1) * B], dm[i * B:(i + 1) * B], genome))
    torch.cuda.synchronize()
    L.mural_profile_begin()
    for i in range(3, 7):
        ts.step(SiteBatch(dp[i * B:(i + 1) * B], dm[i * B:(i + 1) * B], genome))
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    L.mural_profile_end(buf, len(buf))
    prof = json.loads(buf.value.decode())
    tot = sum(v["ms"] for v in prof.values()); n = sum(v["count"] for v in prof.values())
    print("B=%d: %d launches / 4 steps, %.3f ms kernel time per step" % (B, n, tot / 4))
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]:
        print("   %-40s n=%4d  %8.3f ms/step" % (k[:40], v["count"] // 4, v["ms"] / 4))
