"""bf16 path vs fp32 kernels (which match the reference to 1e-6) on many sites, every shipped-checkpoint fixture."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import bench
from conftest import SNV_TAGS, load_snv_golden
from test_gpu_snv_forward import build_model
from mural_b200 import PackedGenome, SiteBatch, pack_meta
chrom = bench.synth_chromosome(0)
genome = PackedGenome({"chr1": chrom.tobytes()})
rng = np.random.default_rng(7)
n = 400000
st = np.sort(rng.choice(np.arange(15000, 2_000_000), n, replace=False)).astype(np.int32)
sd = rng.integers(0, 2, n)
sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), genome)
isN = (chrom == ord("N")).astype(np.int64)
cs = np.r_[0, np.cumsum(isN)]
def has_n(R):
    return (cs[st + R + 1] - cs[st - R]) > 0
for tag in SNV_TAGS:
    z, cfg, state = load_snv_golden(tag)
    m = build_model(cfg, state, int(z["n_cat"]), mode="fp32")
    with torch.no_grad():
        a = torch.softmax(m.forward(None, sb), 1)
        m.compute_mode = "bf16"
        b = torch.softmax(m.forward(None, sb), 1)
    d = (a - b).abs().max(1).values
    q = torch.quantile(d.float(), torch.tensor([0.5, 0.99, 0.9999], device=d.device)).tolist()
    hn = torch.from_numpy(has_n(cfg["distal_radius"])).cuda()
    print("%-10s R_d=%4d  max|dp| = %.3e   median %.1e  p99 %.1e  p99.99 %.1e   sites > 5e-3: %d / %d | windows with N: %d, max|dp| there %.3e, elsewhere %.3e" %
          (tag, cfg["distal_radius"], d.max().item(), q[0], q[1], q[2], int((d > 5e-3).sum()), n, int(hn.sum()), d[hn].max().item() if hn.any() else 0.0, d[~hn].max().item()))
