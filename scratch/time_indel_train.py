"""Wall-clock (CUDA events on the step stream) of the MuRaL-indel training step at batch 32, L = 8000."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from mural_b200 import PackedGenome, SiteBatch, model_choice, pack_meta
from mural_b200.training import IndelTrainState
chroms = [bench.synth_chromosome(0)]
genome = PackedGenome({"chr1": chroms[0].tobytes()})
z = np.load(os.path.join(ROOT, "tests", "golden", "indel_hs_ins.npz"))
state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
cfg = {"CNN_out_channels": 8, "CNN_kernel_size": 7, "down_list": [int(v) for v in z["down"]], "use_reverse": bool(z["use_reverse"]), "n_class": 8}
m = model_choice(0, cfg, {"n_class": 8}, "indel")
m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
m.to("cuda").train()
ts = IndelTrainState(m, 4000, "Adam", lr=1e-4, weight_decay=1e-5, seed=0)
for B in (32, 128):
    n = B * 24
    rng = np.random.default_rng(1)
    lab = rng.choice(8, size=n, p=[0.907] + [0.093 / 7] * 7)
    pos = torch.from_numpy((20000 + 50 * np.arange(n)).astype(np.int32)).cuda()
    meta = torch.from_numpy(pack_meta(np.zeros(n, np.int64), lab, np.zeros(n, np.int64))).cuda()
    for i in range(4):
        ts.step(SiteBatch(pos[i * B:(i + 1) * B], meta[i * B:(i + 1) * B], genome))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(4, 24):
        ts.step(SiteBatch(pos[i * B:(i + 1) * B], meta[i * B:(i + 1) * B], genome))
    e1.record(); torch.cuda.synchronize()
    print("batch %d: %.3f ms per step, loss finite %s" % (B, e0.elapsed_time(e1) / 20, bool(np.isfinite(float(ts.loss_dev.item())))))
