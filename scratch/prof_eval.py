"""ncu target: the validation-metric reductions on 1M synthetic validation sites (k=3,5,7 tables, regional tables, window runs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mural_b200.evaluation import EvalData, kmer_group_table, window_table
rng = np.random.default_rng(0)
n, K = 1_000_000, 4
flank = torch.from_numpy(rng.integers(0, 4, (n, 15)).astype(np.int64)).cuda()
labels = rng.choice(4, n, p=[0.952381, 0.0140095, 0.0198, 0.0138095])
start = np.sort(rng.integers(0, 100_000_000, n)).astype(np.int32)
meta = torch.from_numpy((labels << 1).astype(np.int32)).cuda()
prob = torch.from_numpy(rng.dirichlet(np.ones(K), n)).cuda()
ed = EvalData(flank, meta, prob, start=torch.from_numpy(start).cuda())
for k in (3, 5, 7):
    kmer_group_table(ed, k)
kmer_group_table(ed, 5, 10000)
window_table(ed, 100000)
torch.cuda.synchronize()
print("done")
