"""GPU: Network2 with continuous features (n_cont = 2: first_bn_layer + the wider first Linear of the local branch,
model_snv.py:326-334,457-463) vs the UNMODIFIED reference module's log-probs (oracle/make_golden_cont.py), on the site-record
fast path and on the reference's tensor signature."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLD

pytestmark = pytest.mark.gpu


def test_network2_with_continuous_features(kat, cuda_genome):
    from mural_b200 import SiteBatch, model_choice, pack_meta
    z = np.load(os.path.join(GOLD, "snv_cont_kat.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    n_cat, n_cont = int(z["n_cat"]), int(z["n_cont"])
    common = dict(emb_dims=[(65, 2)] * n_cat, n_cont=n_cont, n_class=4, distal_order=1, in_channels=4)
    m = model_choice(2, cfg, common, "snv")
    sd = m.state_dict()
    assert sd["first_bn_layer.weight"].shape == (n_cont,) and sd["lin_layers.0.weight"].shape == (150, 5 * n_cat + n_cont)
    alias = {"1": "bn1", "2": "conv1", "4": "bn2", "5": "conv2"}
    for k in sd:
        src = k
        if ".layer." in k:
            head, rest = k.split(".layer.")
            src = head + "." + alias[rest.split(".")[0]] + "." + rest.split(".", 1)[1]
        sd[k] = torch.from_numpy(np.asarray(state[src]))
    m.load_state_dict(sd, strict=True)
    m.to("cuda").eval()
    pos = torch.from_numpy(z["start"].astype(np.int32)).cuda()
    meta = torch.from_numpy(pack_meta(z["strand"], 0 * z["strand"], z["chrom"])).cuda()
    cont = torch.from_numpy(z["cont"]).cuda()
    ref = z["ref_logp"]
    with torch.no_grad():
        a = m.forward(None, SiteBatch(pos, meta, cuda_genome, cont=cont)).cpu().numpy()
        cat = cuda_genome.encode_local(pos, meta, cfg["local_radius"], cfg["local_order"])
        oh = cuda_genome.encode_onehot(pos, meta, cfg["distal_radius"])
        b = m.forward((cont, cat), oh).cpu().numpy()
        m.compute_mode = "auto"                                   # continuous-feature models stay on the fp32-equivalent path
        c = m.forward(None, SiteBatch(pos, meta, cuda_genome, cont=cont)).cpu().numpy()
    sm = lambda x: torch.softmax(torch.from_numpy(x), 1).numpy()
    assert np.abs(sm(a) - sm(ref)).max() <= 1e-3 and np.abs(a - ref).max() < 1e-4
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # the features matter: other values move the output
    with torch.no_grad():
        d = m.forward(None, SiteBatch(pos, meta, cuda_genome, cont=cont * 0)).cpu().numpy()
    assert np.abs(d - a).max() > 1e-3
    with pytest.raises(RuntimeError):
        m.forward(None, SiteBatch(pos, meta, cuda_genome))                     # cont_x missing
    # ---- train mode: first_bn_layer on batch statistics, the wider first Linear, gradients vs fp64 autograd of the oracle.
    # (The reference's numpy batch pipeline hands cont_x = zeros to its training loop, Create_DatasetSegment, preprocessing.py:1209;
    # the training kernels take the real features through SiteBatch.cont.)
    from mural_b200 import _lib
    from mural_b200.training import TrainState
    from oracle import encode_np as E
    from oracle import network_t as NT
    n = 40
    labels = (z["start"][:n] % 4).astype(np.int64)
    st = TrainState(m, "Adam", lr=1e-3)
    st.set_dropout(0, 0, 0)
    m.train()
    sb = SiteBatch(pos[:n], torch.from_numpy(pack_meta(z["strand"][:n], labels, z["chrom"][:n])).cuda(), cuda_genome, cont=cont[:n])
    logp = st.forward(sb)
    _, genome = kat
    names = list(genome)
    cat_o = np.empty((n, n_cat), np.int64)
    oh_o = np.empty((n, 4, 2 * cfg["distal_radius"] + 1), np.float32)
    for ci in range(len(names)):
        msk = z["chrom"][:n] == ci
        if msk.any():
            sym = E.seq_to_symbols(genome[names[ci]])
            cat_o[msk] = E.kmer_windows(sym, z["start"][:n][msk], z["strand"][:n][msk], cfg["local_radius"], cfg["local_order"])
            oh_o[msk] = E.onehot_windows(sym, z["start"][:n][msk], z["strand"][:n][msk], cfg["distal_radius"])
    sd64 = {k: torch.tensor(np.asarray(v), dtype=torch.float64, requires_grad=("running" not in k)) for k, v in state.items()
            if "num_batches" not in k}
    refo = NT.network2_forward(sd64, cat_o, oh_o, torch.float64, train=True, cont_x=z["cont"][:n].astype(np.float64))
    assert np.abs(logp.cpu().numpy() - refo.detach().numpy()).max() < 2e-4
    NT.ce_sum(refo, labels).backward()
    dlogp = torch.empty_like(logp)
    _lib.check(_lib.lib().mural_ce_sum_grad(_lib.ptr(logp), _lib.ptr(sb.meta), n, 4, _lib.ptr(st.loss_dev), _lib.ptr(dlogp), _lib.current_stream()))
    g = st.backward(dlogp).cpu().numpy()
    for name, off, num, is_buf in m.native_layout():
        if is_buf or not (name.startswith("first_bn_layer") or name.startswith("lin_layers.0") or name.startswith("emb_layer")):
            continue
        ref_g = sd64[name].grad.numpy().reshape(-1)
        err = np.abs(g[off:off + num] - ref_g).max() / max(1e-6, np.abs(ref_g).max())
        assert err < 5e-3, (name, err)
    # two fused steps (graph capture on the second) run and keep the loss finite
    for _ in range(3):
        st.step(sb)
    assert np.isfinite(float(st.loss_dev.item()))
