"""GPU: UNet_Small eval forward vs the reference outputs (fixtures) for three shipped INDEL checkpoints."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD, INDEL_TAGS

pytestmark = pytest.mark.gpu


def _model(z):
    from mural_b200 import model_choice
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    cfg = {"CNN_out_channels": state["uplblocks.0.0.weight"].shape[0], "CNN_kernel_size": state["uplblocks.0.0.weight"].shape[2],
           "down_list": [int(v) for v in z["down"]], "use_reverse": bool(z["use_reverse"]), "n_class": state["out_fc.2.weight"].shape[0]}
    m = model_choice(0, cfg, {"n_class": cfg["n_class"]}, "indel")
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(state.keys())
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    return m.to("cuda").eval(), state


@pytest.mark.parametrize("tag", INDEL_TAGS)
def test_indel_forward_matches_reference(kat, cuda_genome, tag, manifest):
    from mural_b200 import SiteBatch, pack_meta
    z = np.load(os.path.join(GOLD, "indel_%s.npz" % tag))
    m, state = _model(z)
    assert [k for k, _, _ in manifest["state_dict_keys"][tag]] == list(m.state_dict().keys())
    Rd = int(z["distal_radius"])
    pos = torch.from_numpy(z["start"].astype(np.int32)).cuda()
    meta = torch.from_numpy(pack_meta(z["strand"], 0 * z["strand"], z["chrom"])).cuda()
    sb = SiteBatch(pos, meta, cuda_genome)
    with torch.no_grad():
        a = m.forward(sb, distal_radius=Rd)
        oh = cuda_genome.encode_onehot(pos, meta, Rd, "indel")
        b = m.forward(oh)
    assert torch.equal(a, b)                                   # site path == the reference's tensor signature
    ref = z["ref_out"]
    d = np.abs(a.cpu().numpy() - ref).max()
    print(tag, "max |out - ref| = %.2e (scale %.2f)" % (d, np.abs(ref).max()))
    assert d <= 1e-3 * max(1.0, np.abs(ref).max())            # fp32-equivalent gate
    # train mode routes to the training tape (tests/test_gpu_indel_train.py pins its values): differentiable output
    m.train()
    out = m.forward(oh[:4])
    assert out.shape == (4, a.shape[1]) and out.grad_fn is not None and bool(torch.isfinite(out).all())
