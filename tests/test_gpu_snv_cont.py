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
    from mural_b200.training import TrainState
    with pytest.raises(NotImplementedError):
        TrainState(m)
