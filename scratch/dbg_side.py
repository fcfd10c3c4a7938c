import os, sys, numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from mural_b200 import SiteBatch, model_choice, pack_meta, weights_init
from conftest import *  # noqa
import mural_b200
from mural_b200 import PackedGenome
R_d, R_l = int(sys.argv[1]), 10
torch.manual_seed(R_d)
cfg = {"local_radius": R_l, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": R_d,
       "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": 3, "CNN_out_channels": 32, "distal_fc_dropout": .25,
       "n_class": 4, "model_no": 2}
common = dict(emb_dims=[(65, 2)] * (2 * R_l - 1), n_cont=0, n_class=4, distal_order=1, in_channels=4)
m = model_choice(2, cfg, common, "snv")
m.apply(weights_init)
for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm1d) and mod.num_features > 0:
        mod.running_mean.normal_(0, .3); mod.running_var.uniform_(.5, 1.5)
        mod.weight.data.uniform_(.5, 1.5); mod.bias.data.normal_(0, .2)
m.to("cuda").eval()
rng = np.random.default_rng(R_d)
n = 3000
z = np.load(os.path.join(GOLD, "encode_kat.npz"))
g = PackedGenome({str(a): str(b) for a, b in zip(z["genome_names"], z["genome_seqs"])})
st = np.sort(rng.integers(0, 30000, n)).astype(np.int32)
sd = rng.integers(0, 2, n)
sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), g)
res = {}
if len(sys.argv) > 2:
    with torch.no_grad():
        m.compute_mode = "bf16"
        r = m.forward(None, sb).clone()
    torch.cuda.synchronize()
    print("only lattice run", float(r.abs().sum()))
    sys.exit(0)
with torch.no_grad():
    m.compute_mode = "fp32"
    res["fp32"] = m.forward(None, sb).clone()
    m.compute_mode = "bf16"
    for key, env in (("lattice", None), ("site", "1"), ("lattice2", None)):
        if env is None:
            os.environ.pop("MURAL_NO_LATTICE", None)
        else:
            os.environ["MURAL_NO_LATTICE"] = env
        res[key] = m.forward(None, sb).clone()
sm = lambda k: torch.softmax(res[k], 1)
print("R", R_d, "lat-site", float((res["lattice"] - res["site"]).abs().max()), "lat-lat2", float((res["lattice"] - res["lattice2"]).abs().max()),
      "lat-fp32", float((sm("lattice") - sm("fp32")).abs().max()), "site-fp32", float((sm("site") - sm("fp32")).abs().max()))
