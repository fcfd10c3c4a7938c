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
    # the shipped shapes run the fused tensor-core level kernels by default; the fp32 CUDA-core kernels hold the same gate
    from mural_b200 import _lib
    assert _lib.lib().mural_indel_tc_available(m._handle(Rd)) == 1
    m.compute_mode = "fp32"
    with torch.no_grad():
        c = m.forward(sb, distal_radius=Rd)
    m.compute_mode = "auto"
    d32 = np.abs(c.cpu().numpy() - ref).max()
    dtc = (a - c).abs().max().item()
    print(tag, "fp32 kernels: max |out - ref| = %.2e; tensor-core vs fp32 kernels %.2e" % (d32, dtc))
    assert d32 <= 1e-3 * max(1.0, np.abs(ref).max()) and dtc <= 1e-3 * max(1.0, np.abs(ref).max())
    # train mode routes to the training tape (tests/test_gpu_indel_train.py pins its values): differentiable output
    m.train()
    out = m.forward(oh[:4])
    assert out.shape == (4, a.shape[1]) and out.grad_fn is not None and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("C,ks,down,R,rev", [(8, 7, [1, 4, 5, 5, 5, 2], 500, True), (8, 7, [1, 4, 5, 5, 5, 2], 1500, False),
                                             (8, 5, [2, 2, 2, 2, 2, 2], 544, True), (8, 3, [1, 2, 3, 2, 1, 2], 1260, False),
                                             (8, 9, [4, 1, 5, 2, 2, 1], 3000, True)])
def test_indel_level_kernels_generic_shapes(C, ks, down, R, rev):
    """Fused tensor-core level kernels vs the fp32 kernels and the CPU oracle on random weights: tile edges (lengths that are
    not multiples of the tile), strides / upsampling factors other than the shipped ones, both stems."""
    from mural_b200 import _lib, model_choice
    from oracle import network_t as NT
    torch.manual_seed(ks * 100 + R)
    cfg = {"CNN_out_channels": C, "CNN_kernel_size": ks, "down_list": down, "use_reverse": rev, "n_class": 8}
    m = model_choice(0, cfg, {"n_class": 8}, "indel")
    for mod in m.modules():                                  # non-trivial BatchNorm statistics
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.running_mean.normal_(0, 0.3); mod.running_var.uniform_(0.5, 1.5); mod.weight.data.uniform_(0.5, 1.5); mod.bias.data.normal_(0, 0.2)
    m.to("cuda").eval()
    n = 5
    idx = torch.randint(0, 4, (n, 2 * R))
    oh = torch.nn.functional.one_hot(idx, 4).permute(0, 2, 1).float()
    oh[0, :, 10:40] = 0.25                                   # an N run
    with torch.no_grad():
        a = m.forward(oh.cuda())
        assert _lib.lib().mural_indel_tc_available(m._handle(R)) == 1
        m.compute_mode = "fp32"
        b = m.forward(oh.cuda())
        sd = {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        ref = NT.unet_small_forward(sd, oh.numpy(), down, rev, torch.float64).numpy()
    scale = max(1.0, np.abs(ref).max())
    da, db = np.abs(a.cpu().numpy() - ref).max(), np.abs(b.cpu().numpy() - ref).max()
    print("C=%d ks=%d down=%s R=%d: tensor-core %.2e, fp32 kernels %.2e (scale %.2f)" % (C, ks, down, R, da, db, scale))
    assert da <= 1e-3 * scale and db <= 1e-3 * scale


def test_indel_level_kernels_batch_edges(kat, cuda_genome):
    """Tensor-core level kernels on awkward batch sizes: 1-3 sites, a count that is not a multiple of any site group, more sites
    than one workspace chunk (4096) — equal to the fp32 kernels within the gate, identical from run to run (the position max is
    a float atomic max: order-independent), and rows do not depend on what else is in the batch."""
    from mural_b200 import SiteBatch, pack_meta
    z = np.load(os.path.join(GOLD, "indel_at_ins.npz"))          # R = 2000: a 4100-site batch stays small
    m, _ = _model(z)
    Rd = int(z["distal_radius"])
    _, genome = kat
    names = list(genome)
    rng = np.random.default_rng(5)
    n = 4100
    ch = rng.integers(0, len(names), n)
    st = np.array([rng.integers(0, len(genome[names[c]])) for c in ch])      # windows may overhang the chromosome ends (N imputation)
    sd = rng.integers(0, 2, n)
    pos = torch.from_numpy(st.astype(np.int32)).cuda()
    meta = torch.from_numpy(pack_meta(sd, 0 * sd, ch)).cuda()
    with torch.no_grad():
        full = m.forward(SiteBatch(pos, meta, cuda_genome), distal_radius=Rd)
        again = m.forward(SiteBatch(pos, meta, cuda_genome), distal_radius=Rd)
        assert torch.equal(full, again)
        for k in (1, 2, 3, 37):
            part = m.forward(SiteBatch(pos[:k], meta[:k], cuda_genome), distal_radius=Rd)
            assert torch.equal(part, full[:k]), k
        tail = m.forward(SiteBatch(pos[4090:], meta[4090:], cuda_genome), distal_radius=Rd)
        assert torch.equal(tail, full[4090:])
        m.compute_mode = "fp32"
        ref = m.forward(SiteBatch(pos[:512], meta[:512], cuda_genome), distal_radius=Rd)
        ref_tail = m.forward(SiteBatch(pos[4000:], meta[4000:], cuda_genome), distal_radius=Rd)
    scale = max(1.0, float(ref.abs().max()))
    assert float((full[:512] - ref).abs().max()) <= 1e-3 * scale
    assert float((full[4000:] - ref_tail).abs().max()) <= 1e-3 * scale
    assert bool(torch.isfinite(full).all())
