"""Per-tensor gradient errors of the training backward vs fp64 autograd (scratch diagnostic)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import GOLD, load_snv_golden
from oracle import network_t as NT
from test_gpu_snv_forward import build_model
from test_gpu_snv_train import _batch, _oracle_inputs
from mural_b200 import PackedGenome, _lib
from mural_b200.training import TrainState
tag = sys.argv[1] if len(sys.argv) > 1 else "ex_ckpt6"
k = np.load(os.path.join(GOLD, "encode_kat.npz"))
genome = {str(n): str(s) for n, s in zip(k["genome_names"], k["genome_seqs"])}
pg = PackedGenome(genome)
z, cfg, state = load_snv_golden(tag)
n = 48
labels = (z["start"][:n] % 4).astype(np.int64)
m = build_model(cfg, state, int(z["n_cat"]))
st = TrainState(m, "Adam", lr=1e-3); st.set_dropout(0, 0, 0); m.train()
sb = _batch(z, pg, n, labels)
logp = st.forward(sb)
cat, oh = _oracle_inputs(z, cfg, genome, n)
sd64 = {kk: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=("running" not in kk)) for kk, v in state.items() if "num_batches" not in kk}
ref = NT.network2_forward(sd64, cat, oh, torch.float64, train=True)
print("logp err", np.abs(logp.cpu().numpy() - ref.detach().numpy()).max())
NT.ce_sum(ref, labels).backward()
dlogp = torch.empty_like(logp)
_lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(sb.meta), n, 4, _lib.ptr(st.loss_dev), _lib.ptr(dlogp), _lib.current_stream()))
g = st.backward(dlogp).cpu().numpy()
rows = []
for name, off, num, is_buf in m.native_layout():
    if is_buf: continue
    r = sd64[name].grad.numpy().reshape(-1)
    rows.append((np.abs(g[off:off+num] - r).max() / max(1e-3, np.abs(r).max()), name, np.abs(r).max()))
rows.sort(reverse=True)
for e, nm, sc in rows[:8]:
    print("%.3e  %-28s |g|max %.3e" % (e, nm, sc))
if len(sys.argv) > 2:
    nm = sys.argv[2]
    for name, off, num, is_buf in m.native_layout():
        if name == nm:
            r = sd64[name].grad.numpy().reshape(-1)
            d = g[off:off+num] - r
            print(nm, "abs err per element:", np.array2string(np.abs(d), precision=2, max_line_width=200))
            print("ref:", np.array2string(r, precision=3, max_line_width=200))
    # batch statistics of the BN in front of RBs1.1.conv1 (input relu(y1))
