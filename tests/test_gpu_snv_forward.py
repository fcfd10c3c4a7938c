"""GPU: Network2 eval forward through the C ABI vs the reference logits (fixtures) and the oracle.

Tolerances (BASELINE.json north_star): |p - p_ref| <= 1e-3 in fp32-equivalent mode, <= 5e-3 in bf16 mode,
on per-site class probabilities.  The fp32 path is also held to a much tighter 2e-5 on log-probs.
"""
import numpy as np
import pytest
import torch

from conftest import SNV_TAGS, load_snv_golden
from oracle import encode_np as E
from oracle import network_t as NT

pytestmark = pytest.mark.gpu


def build_model(cfg, state, n_cat, mode="fp32"):
    from mural_b200 import model_choice
    common = dict(emb_dims=[(65, 2)] * n_cat, n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
    m = model_choice(2, cfg, common, "snv")
    sd = m.state_dict()
    for k in sd:
        src = k
        if ".layer." in k:      # aliases: fill from the canonical tensor
            head, idx, leaf = k.split(".layer.")[0], k.split(".layer.")[1].split(".")[0], k.split(".")[-1]
            src = head + "." + {"1": "bn1", "2": "conv1", "4": "bn2", "5": "conv2"}[idx] + "." + leaf
        sd[k] = torch.from_numpy(np.asarray(state[src]))
    m.load_state_dict(sd, strict=True)
    m.to("cuda").eval()
    m.compute_mode = mode
    return m


def site_batch(z, genome_dev):
    from mural_b200 import SiteBatch, pack_meta
    pos = torch.from_numpy(z["start"].astype(np.int32)).cuda()
    meta = torch.from_numpy(pack_meta(z["strand"], np.zeros(len(z["start"]), int), z["chrom"])).cuda()
    return SiteBatch(pos, meta, genome_dev)


def probs(lp):
    return torch.softmax(torch.as_tensor(lp), 1).numpy()


@pytest.mark.parametrize("tag", SNV_TAGS)
def test_fp32_forward_matches_reference(kat, cuda_genome, tag):
    z, cfg, state = load_snv_golden(tag)
    m = build_model(cfg, state, int(z["n_cat"]))
    with torch.no_grad():
        lp = m.forward(None, site_batch(z, cuda_genome)).cpu().numpy()
    assert lp.shape == z["ref_logp"].shape
    assert np.abs(probs(lp) - probs(z["ref_logp"])).max() <= 1e-3          # the north-star gate
    assert np.abs(lp - z["ref_logp"]).max() < 5e-5                          # fp32 path is far tighter


@pytest.mark.parametrize("tag", ["hs_AT", "mm_CpG", "ex_ckpt6"])
def test_site_chain_equals_per_layer_path(kat, cuda_genome, tag):
    """The fused per-site chain of a CNN branch (snv_site_chain.cu: all 10 convs + 2 pools of a branch in shared memory) runs the
    arithmetic of the per-layer fp32-equivalent kernels in the same order: log-probs equal bit for bit (MURAL_NO_SITE_CHAIN=1
    keeps the per-layer path), for one chunk and for sites spread over several chunks."""
    import os
    z, cfg, state = load_snv_golden(tag)
    m = build_model(cfg, state, int(z["n_cat"]))
    sb = site_batch(z, cuda_genome)
    res = {}
    try:
        for key, env, chunk in (("chain", None, 0), ("layers", "1", 0), ("chain_chunks", None, 50)):
            if env is None:
                os.environ.pop("MURAL_NO_SITE_CHAIN", None)
            else:
                os.environ["MURAL_NO_SITE_CHAIN"] = env
            m.set_debug(False, chunk=chunk)
            with torch.no_grad():
                res[key] = m.forward(None, sb).clone()
    finally:
        os.environ.pop("MURAL_NO_SITE_CHAIN", None)
        m.set_debug(False, chunk=0)
    assert torch.isfinite(res["chain"]).all()
    assert torch.equal(res["chain"], res["layers"]), float((res["chain"] - res["layers"]).abs().max())
    assert torch.equal(res["chain"], res["chain_chunks"])


def test_fp32_intermediate_taps(kat, cuda_genome):
    z, cfg, state = load_snv_golden("hs_AT")
    m = build_model(cfg, state, int(z["n_cat"]))
    m.set_debug(True, chunk=0)
    with torch.no_grad():
        m.forward(None, site_batch(z, cuda_genome))
    C = cfg["CNN_out_channels"]
    for name in ("pool1", "pool1_2", "rb1_2", "conv2_2", "rb2_2"):
        ref = z["tap_" + name]                       # [16, C, L] (torch layout) for the first 16 sites
        got = m.debug_tap(name).reshape(len(z["start"]), -1, C)[:16].transpose(0, 2, 1)
        assert got.shape == ref.shape, name
        # fp32-equivalent: the convs run as split-bf16 tensor-core MMAs (16 mantissa bits per operand, ~1e-5 per product)
        assert np.abs(got - ref).max() < 1e-5 + 1e-5 * np.abs(ref).max(), name      # scale-relative
    for name in ("gmax", "gmax_2", "logit_local", "logit_mid", "logit_large"):
        ref = z["tap_" + name]
        got = m.debug_tap(name).reshape(len(z["start"]), -1)[:16]
        assert np.abs(got - ref).max() < 1e-4 + 1e-5 * np.abs(ref).max(), name
    m.set_debug(False)


def test_tensor_path_and_chunking(kat, cuda_genome):
    """Drop-in signature forward((cont, cat), distal_onehot) == site path; chunk size does not matter."""
    z, cfg, state = load_snv_golden("ex_ckpt6")
    m = build_model(cfg, state, int(z["n_cat"]))
    sb = site_batch(z, cuda_genome)
    with torch.no_grad():
        a = m.forward(None, sb)
        cat = cuda_genome.encode_local(sb.pos, sb.meta, cfg["local_radius"], cfg["local_order"])
        oh = cuda_genome.encode_onehot(sb.pos, sb.meta, cfg["distal_radius"])
        b = m.forward((torch.zeros(len(sb), 1, device="cuda"), cat), oh)
        m.set_debug(False, chunk=37)
        c = m.forward(None, sb)
        m.set_debug(False, chunk=0)
    assert torch.equal(a, b) and torch.equal(a, c)
    with pytest.raises(RuntimeError):
        m.forward((None, cat), oh * 0.7)            # not a reference one-hot encoding
    bad = cat.clone(); bad[0, 0] = 99
    with pytest.raises(IndexError):
        m.forward((None, bad), oh)
    with pytest.raises(AssertionError):
        m.forward((None, cat), oh[:, :, :150])


def test_host_e2e_and_predict_loop(kat, cuda_genome):
    from mural_b200 import PackedSiteDataset, SiteTable, generate_site_batches, model_predict_m, pack_meta
    z, cfg, state = load_snv_golden("hs_AT")
    _, genome = kat
    m = build_model(cfg, state, int(z["n_cat"]))
    meta = pack_meta(z["strand"], z["start"] % 4, z["chrom"])
    out = m.predict_host(cuda_genome, z["start"].astype(np.int32), meta)
    assert np.abs(out - z["ref_logp"]).max() < 5e-5
    # model_predict_m over the dataset/batcher == one big call, loss == CE(sum) of the oracle
    t = SiteTable(list(genome), z["chrom"], z["start"], z["start"] + 1, z["strand"], z["start"] % 4)
    ds = PackedSiteDataset(t, cuda_genome, 300000, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"])
    pred, loss = model_predict_m(m, generate_site_batches(ds, 1, 16), None, torch.device("cuda"), 4)
    inv = np.argsort(ds.perm)
    assert np.abs(pred.cpu().numpy()[inv] - z["ref_logp"]).max() < 5e-5
    exp = float(NT.ce_sum(torch.from_numpy(z["ref_logp"]), z["start"] % 4))
    assert abs(loss - exp) < 1e-3 * max(1.0, abs(exp))


def test_window_sweep_and_local10_vs_oracle(cuda_genome, kat):
    """Random-init weights: other window sizes (config 5 sweep) and the 'local 10 bp' variant vs the oracle."""
    from mural_b200 import SiteBatch, model_choice, pack_meta, weights_init
    _, genome = kat
    names = list(genome)
    torch.manual_seed(0)
    rng = np.random.default_rng(9)
    for R_d, R_l, ks in ((100, 7, 3), (500, 10, 3), (2000, 10, 3), (5000, 7, 3), (300, 7, 5)):
        cfg = {"local_radius": R_l, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": R_d,
               "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": ks, "CNN_out_channels": 32, "distal_fc_dropout": .25,
               "n_class": 4, "model_no": 2}
        n_cat = 2 * R_l + 1 - 2
        common = dict(emb_dims=[(65, 2)] * n_cat, n_cont=0, n_class=4, distal_order=1, in_channels=4)
        m = model_choice(2, cfg, common, "snv")
        m.apply(weights_init)
        for mod in m.modules():                       # non-trivial BN statistics
            if isinstance(mod, torch.nn.BatchNorm1d) and mod.num_features > 0:
                mod.running_mean.normal_(0, .3); mod.running_var.uniform_(.5, 1.5)
                mod.weight.data.uniform_(.5, 1.5); mod.bias.data.normal_(0, .2)
        m.to("cuda").eval()
        n = 24
        ch = rng.integers(0, 2, n); st = np.array([rng.integers(0, len(genome[names[c]])) for c in ch]); sd = rng.integers(0, 2, n)
        sb = SiteBatch(torch.from_numpy(st.astype(np.int32)).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, ch)).cuda(), cuda_genome)
        with torch.no_grad():
            got = m.forward(None, sb).cpu().numpy()
        state = {k: v.cpu().numpy() for k, v in m.state_dict().items() if ".layer." not in k}
        cat = np.empty((n, n_cat), np.int64); oh = np.empty((n, 4, 2 * R_d + 1), np.float32)
        for c in range(2):
            msk = ch == c
            sym = E.seq_to_symbols(genome[names[c]])
            cat[msk] = E.kmer_windows(sym, st[msk], sd[msk], R_l, 3); oh[msk] = E.onehot_windows(sym, st[msk], sd[msk], R_d)
        with torch.no_grad():
            exp = NT.network2_forward(state, cat, oh, torch.float64).numpy()
        assert np.abs(probs(got) - probs(exp)).max() <= 1e-3, (R_d, R_l, ks)
        assert np.abs(got - exp).max() < 2e-4, (R_d, R_l, ks)


def test_fast_stem_equals_generic_stem(kat, cuda_genome):
    """4-mer pair-table stem == per-tap table stem, bit for bit (incl. N / IUPAC / chromosome-edge windows)."""
    for tag in ("hs_AT", "ex_ckpt6"):
        z, cfg, state = load_snv_golden(tag)
        m = build_model(cfg, state, int(z["n_cat"]))
        sb = site_batch(z, cuda_genome)
        taps = {}
        for flags in (1, 3):
            m.set_debug(flags)
            with torch.no_grad():
                out = m.forward(None, sb)
            taps[flags] = (m.debug_tap("pool1").copy(), m.debug_tap("pool1_2").copy(), out.clone())
        m.set_debug(0)
        assert np.array_equal(taps[1][0].view(np.uint32), taps[3][0].view(np.uint32))
        assert np.array_equal(taps[1][1].view(np.uint32), taps[3][1].view(np.uint32))
        assert torch.equal(taps[1][2], taps[3][2])
