"""GPU: bf16 tcgen05 path (MURAL_MODE_BF16) of Network2 vs the reference logits; gate 5e-3 on probabilities."""
import numpy as np
import pytest
import torch

from conftest import SNV_TAGS, load_snv_golden
from test_gpu_snv_forward import build_model, probs, site_batch

pytestmark = pytest.mark.gpu


def test_bf16_taps_close_to_reference(kat, cuda_genome):
    """Stage outputs of the chained tcgen05 kernels: relative error at bf16 level, catches layout/descriptor bugs."""
    z, cfg, state = load_snv_golden("hs_AT")
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    m.set_debug(True, chunk=0)
    with torch.no_grad():
        m.forward(None, site_batch(z, cuda_genome))
    C = cfg["CNN_out_channels"]
    report = {}
    for name in ("pool1", "pool1_2", "rb1", "rb1_2", "rb2", "rb2_2"):
        if "tap_" + name not in z.files:
            continue
        ref = z["tap_" + name]
        got = m.debug_tap(name).reshape(len(z["start"]), -1, C)[:16].transpose(0, 2, 1)
        assert got.shape == ref.shape, name
        report[name] = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
    for name in ("gmax", "gmax_2", "logit_local", "logit_mid", "logit_large"):
        ref = z["tap_" + name]
        got = m.debug_tap(name).reshape(len(z["start"]), -1)[:16]
        report[name] = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
    print("bf16 tap relative errors:", report)
    assert report["pool1_2"] < 5e-3 and report["logit_local"] < 1e-4     # stem output is stored as bf16; local branch is fp32
    for k, v in report.items():
        assert v < 3e-2, (k, v, report)
    m.set_debug(False)


@pytest.mark.parametrize("tag", SNV_TAGS)
def test_bf16_forward_matches_reference(kat, cuda_genome, tag):
    z, cfg, state = load_snv_golden(tag)
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    with torch.no_grad():
        lp = m.forward(None, site_batch(z, cuda_genome)).cpu().numpy()
    d = np.abs(probs(lp) - probs(z["ref_logp"])).max()
    print(tag, "bf16 max|dp| = %.3e" % d)
    assert d <= 5e-3


def test_bf16_chunking_and_tensor_path(kat, cuda_genome):
    z, cfg, state = load_snv_golden("hs_nonCpG")
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    sb = site_batch(z, cuda_genome)
    with torch.no_grad():
        a = m.forward(None, sb)
        m.set_debug(False, chunk=29)
        b = m.forward(None, sb)
        m.set_debug(False, chunk=0)
        cat = cuda_genome.encode_local(sb.pos, sb.meta, cfg["local_radius"], cfg["local_order"])
        oh = cuda_genome.encode_onehot(sb.pos, sb.meta, cfg["distal_radius"])
        c = m.forward((None, cat), oh)
    assert torch.equal(a, b), float((a - b).abs().max())     # tiling must not change any row's arithmetic
    assert torch.equal(a, c)


def test_bf16_large_batch_matches_fp32_path(kat, cuda_genome):
    """Many tiles per CTA (persistent loop, phase flips): bf16 vs the fp32 kernels on 20k sites."""
    from mural_b200 import SiteBatch, pack_meta
    z, cfg, state = load_snv_golden("hs_AT")
    _, genome = kat
    rng = np.random.default_rng(4)
    n = 20000
    st = rng.integers(0, 30000, n).astype(np.int32)
    sd = rng.integers(0, 2, n)
    sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), cuda_genome)
    m = build_model(cfg, state, int(z["n_cat"]), mode="fp32")
    with torch.no_grad():
        a = m.forward(None, sb)
        m.compute_mode = "bf16"
        b = m.forward(None, sb)
    d = (torch.softmax(a, 1) - torch.softmax(b, 1)).abs().max().item()
    print("bf16 vs fp32 kernels, 20k sites: max|dp| = %.3e" % d)
    assert d <= 5e-3


def test_dense_site_stem_equals_per_site_stem(kat, cuda_genome):
    """Genome-wide (dense, one chromosome) chunks take the sliding-window stem; its output must be bit-identical to
    the per-site stem, including chromosome-end overhang, N runs, IUPAC codes and both strands."""
    import os
    from mural_b200 import SiteBatch, pack_meta
    z, cfg, state = load_snv_golden("hs_AT")
    rng = np.random.default_rng(12)
    n = 6000
    st = np.sort(rng.integers(0, 30000, n)).astype(np.int32)          # chrA: N runs at both ends, IUPAC block at 7000
    sd = rng.integers(0, 2, n)
    sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), cuda_genome)
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    res = {}
    for key, env in (("dense", None), ("site", "1")):
        if env is None:
            os.environ.pop("MURAL_NO_DENSE_STEM", None)
        else:
            os.environ["MURAL_NO_DENSE_STEM"] = env
        m.set_debug(1)
        with torch.no_grad():
            out = m.forward(None, sb)
        res[key] = (m.debug_tap("pool1").copy(), m.debug_tap("pool1_2").copy(), out.clone())
    os.environ.pop("MURAL_NO_DENSE_STEM", None)
    m.set_debug(0)
    assert np.array_equal(res["dense"][0].view(np.uint32), res["site"][0].view(np.uint32))
    assert np.array_equal(res["dense"][1].view(np.uint32), res["site"][1].view(np.uint32))
    assert torch.equal(res["dense"][2], res["site"][2])


@pytest.mark.parametrize("chunk", [0, 2048])
def test_dense_lattice_equals_per_site_stages(kat, cuda_genome, chunk):
    """Dense chunks evaluate the first ResBlock pair once per genomic position (stage-1 lattice) plus an 18-row edge
    pseudo-site per site; every row takes the same tcgen05 arithmetic as in the per-site row space, so the log-probs
    must be bit-identical to the per-site stages (MURAL_NO_LATTICE=1), across chunk boundaries, chromosome-end
    overhang, N runs, IUPAC codes and both strands."""
    import os
    from mural_b200 import SiteBatch, pack_meta
    z, cfg, state = load_snv_golden("hs_AT")
    rng = np.random.default_rng(13)
    n = 9000
    st = np.sort(rng.integers(0, 30000, n)).astype(np.int32)
    sd = rng.integers(0, 2, n)
    sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), cuda_genome)
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    m.set_debug(False, chunk=chunk)
    res = {}
    for key, env in (("lattice", None), ("site", "1")):
        if env is None:
            os.environ.pop("MURAL_NO_LATTICE", None)
        else:
            os.environ["MURAL_NO_LATTICE"] = env
        with torch.no_grad():
            res[key] = m.forward(None, sb).clone()
    os.environ.pop("MURAL_NO_LATTICE", None)
    m.set_debug(False, chunk=0)
    assert torch.isfinite(res["lattice"]).all()
    bad = (res["lattice"] != res["site"]).any(1).nonzero().flatten()
    assert bad.numel() == 0, (bad[:10].tolist(), st[bad[:10].cpu().numpy()].tolist(),
                              float((res["lattice"] - res["site"]).abs().max()))


@pytest.mark.parametrize("R_d,R_l", [(100, 7), (200, 10), (500, 7), (2000, 10), (5000, 7)])
def test_bf16_window_sweep_dense_sites(kat, cuda_genome, R_d, R_l):
    """Config-5 window sweep on the tcgen05 path with random-init weights and dense sorted sites: the lattice path
    (where the window is long enough for it) equals the per-site stages bit for bit, and both stay within 5e-3 of the
    fp32 kernels."""
    import os
    from mural_b200 import SiteBatch, model_choice, pack_meta, weights_init
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
    st = np.sort(rng.integers(0, 30000, n)).astype(np.int32)
    sd = rng.integers(0, 2, n)
    sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda(), cuda_genome)
    res = {}
    with torch.no_grad():
        m.compute_mode = "fp32"
        res["fp32"] = m.forward(None, sb).clone()
        m.compute_mode = "bf16"
        for key, env in (("lattice", None), ("site", "1")):
            if env is None:
                os.environ.pop("MURAL_NO_LATTICE", None)
            else:
                os.environ["MURAL_NO_LATTICE"] = env
            res[key] = m.forward(None, sb).clone()
    os.environ.pop("MURAL_NO_LATTICE", None)
    assert torch.equal(res["lattice"], res["site"]), float((res["lattice"] - res["site"]).abs().max())
    d = (torch.softmax(res["lattice"], 1) - torch.softmax(res["fp32"], 1)).abs().max().item()
    print("R_d=%d bf16 vs fp32 kernels max|dp| = %.3e" % (R_d, d))
    assert d <= 5e-3


def test_local_branch_tensor_core_equals_fp32_kernel(kat, cuda_genome):
    """Split-bf16 tcgen05 local branch (snv_mlp_tc.cu) vs the fp32 CUDA-core kernel: same folded weights, logits to 1e-5."""
    import os
    z, cfg, state = load_snv_golden("hs_AT")
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    sb = site_batch(z, cuda_genome)
    taps = {}
    for key, env in (("tc", None), ("fp32", "1")):
        if env is None:
            os.environ.pop("MURAL_NO_MLP_TC", None)
        else:
            os.environ["MURAL_NO_MLP_TC"] = env
        m.set_debug(True, chunk=0)
        with torch.no_grad():
            m.forward(None, sb)
        taps[key] = m.debug_tap("logit_local").copy()
    os.environ.pop("MURAL_NO_MLP_TC", None)
    m.set_debug(False)
    d = np.abs(taps["tc"] - taps["fp32"]).max()
    print("local logits tensor-core vs fp32 kernel: max|d| = %.3e (|logit| max %.2f)" % (d, np.abs(taps["fp32"]).max()))
    assert d < 1e-5 * max(1.0, np.abs(taps["fp32"]).max())


def test_bf16_bits_unchanged(kat, cuda_genome):
    """Restructuring the stage kernels (issuer warps, loaders, fusions) must not change a single MMA or rounding: the bf16
    log-probs of fixed dense (lattice / edge / pre-pooled loaders) and sparse (per-site rows, pooled loader) site sets are
    held to the sha256 recorded with the round-1 kernels on a B200 (scratch/make_bf16_bits.py -> tests/golden/bf16_bits.json)."""
    import hashlib
    import json
    import os
    import sys
    from conftest import GOLD, ROOT
    from mural_b200 import SiteBatch, pack_meta
    sys.path.insert(0, os.path.join(ROOT, "scratch"))
    gold = json.load(open(os.path.join(GOLD, "bf16_bits.json")))
    _, genome = kat
    names = list(genome)
    sets = {}
    rng = np.random.default_rng(21)
    n = 7000
    st = np.sort(rng.integers(0, 30000, n)).astype(np.int32); sd = rng.integers(0, 2, n)
    sets["dense"] = (st, sd, np.zeros(n, np.int64))
    ch = rng.integers(0, 2, n)
    st = np.array([rng.integers(0, len(genome[names[c]])) for c in ch]).astype(np.int32); sd = rng.integers(0, 2, n)
    sets["sparse"] = (st, sd, ch)
    for tag in ("hs_AT", "ex_ckpt6"):
        z, cfg, state = load_snv_golden(tag)
        m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
        for name, (st, sd, ch) in sets.items():
            sb = SiteBatch(torch.from_numpy(st).cuda(), torch.from_numpy(pack_meta(sd, 0 * sd, ch)).cuda(), cuda_genome)
            with torch.no_grad():
                lp = m.forward(None, sb).cpu().numpy()
            g = gold["%s/%s" % (tag, name)]
            assert np.array_equal(lp[:4], np.asarray(g["head"], dtype=np.float32)), (tag, name, lp[:4], g["head"])
            assert hashlib.sha256(lp.tobytes()).hexdigest() == g["sha256"], (tag, name)


@pytest.mark.parametrize("n", [0, 1, 2, 127, 129, 257])
def test_bf16_tiny_batches(kat, cuda_genome, n):
    """Empty and tiny batches through the C ABI (tile / chunk / lattice boundaries): same sites give the same rows whatever
    the batch they arrive in, and agree with the fp32 kernels."""
    from mural_b200 import SiteBatch, pack_meta
    z, cfg, state = load_snv_golden("hs_AT")
    rng = np.random.default_rng(100 + n)
    st = np.sort(rng.integers(0, 30000, 300)).astype(np.int32)
    sd = rng.integers(0, 2, 300)
    pos = torch.from_numpy(st).cuda(); meta = torch.from_numpy(pack_meta(sd, 0 * sd, 0 * sd)).cuda()
    m = build_model(cfg, state, int(z["n_cat"]), mode="bf16")
    with torch.no_grad():
        full = m.forward(None, SiteBatch(pos, meta, cuda_genome))
        part = m.forward(None, SiteBatch(pos[:n], meta[:n], cuda_genome))
        assert part.shape == (n, 4)
        if n:
            assert torch.equal(part, full[:n])            # dense-path values depend on the site only, not on its batch
            m.compute_mode = "fp32"
            ref = m.forward(None, SiteBatch(pos[:n], meta[:n], cuda_genome))
            assert (torch.softmax(part, 1) - torch.softmax(ref, 1)).abs().max().item() <= 5e-3
