"""CPU: the oracle (oracle/) against the fixtures generated from the real reference."""
import os
import numpy as np
import pytest
import torch

from conftest import INDEL_TAGS, SNV_TAGS, load_snv_golden, GOLD
from oracle import encode_np as E
from oracle import network_t as NT

CASES = range(6)


def _oracle_case(z, genome, ci):
    mt, R_l, order, R_d, central = [int(v) for v in z["case%d_cfg" % ci]]
    mt = "snv" if mt == 0 else "indel"
    names = list(genome)
    ch, st, sd = z["chrom"], z["start"], z["strand"]
    perm, sizes = E.order_sites(ch, st, sd, central)
    syms = [E.seq_to_symbols(genome[n]) for n in names]
    cat = np.empty((len(perm), E.window_length(R_l, mt) - (order - 1)), dtype=np.int64)
    oh = np.empty((len(perm), 4, E.window_length(R_d, mt)), dtype=np.float32)
    for c in range(len(names)):
        m = ch[perm] == c
        cat[m] = E.kmer_windows(syms[c], st[perm][m], sd[perm][m], R_l, order, mt)
        oh[m] = E.onehot_windows(syms[c], st[perm][m], sd[perm][m], R_d, mt)
    return perm, sizes, cat, oh


@pytest.mark.parametrize("ci", CASES)
def test_encoders_match_reference_fixture(kat, ci):
    z, genome = kat
    perm, sizes, cat, oh = _oracle_case(z, genome, ci)
    assert np.array_equal(perm, z["case%d_perm" % ci])
    assert np.array_equal(sizes, z["case%d_sizes" % ci])
    assert np.array_equal(cat, z["case%d_cat" % ci])                      # integer: bit exact
    rows = z["case%d_oh_rows" % ci]
    assert np.array_equal(oh[rows].view(np.uint32), z["case%d_oh" % ci].view(np.uint32))   # fp32 bit patterns


def test_literal_bed_reader_equals_vectorised(kat):
    z, _ = kat
    rng = np.random.default_rng(0)
    for central in (1, 7, 333, 5000, 10 ** 6):
        lit = E.bed_batches(z["chrom"], z["start"], z["strand"], central)
        perm, sizes = E.order_sites(z["chrom"], z["start"], z["strand"], central)
        assert np.array_equal(np.concatenate([np.array(b[0]) for b in lit]), perm)
        assert [len(b[0]) for b in lit] == list(sizes)
    # unsorted / interleaved chromosomes: the state machine never moves a window backwards
    ch = rng.integers(0, 3, 500); st = rng.integers(0, 50000, 500); sd = rng.integers(0, 2, 500)
    lit = E.bed_batches(ch, st, sd, 4000)
    perm, sizes = E.order_sites(ch, st, sd, 4000)
    assert np.array_equal(np.concatenate([np.array(b[0]) for b in lit]), perm)


def test_unknown_character_is_keyerror():
    with pytest.raises(KeyError):
        E.seq_to_symbols("ACGTX")


@pytest.mark.parametrize("tag", SNV_TAGS)
def test_network2_oracle_matches_reference_logits(kat, tag, manifest):
    z, cfg, state = load_snv_golden(tag)
    _, genome = kat
    names = list(genome)
    syms = [E.seq_to_symbols(genome[n]) for n in names]
    ch, st, sd = z["chrom"], z["start"], z["strand"]
    n = len(st)
    cat = np.empty((n, int(z["n_cat"])), dtype=np.int64)
    oh = np.empty((n, 4, 2 * cfg["distal_radius"] + 1), dtype=np.float32)
    for c in range(len(names)):
        m = ch == c
        cat[m] = E.kmer_windows(syms[c], st[m], sd[m], cfg["local_radius"], cfg["local_order"])
        oh[m] = E.onehot_windows(syms[c], st[m], sd[m], cfg["distal_radius"])
    with torch.no_grad():
        taps = {}
        lp = NT.network2_forward(state, cat, oh, torch.float32, taps=taps).numpy()
    assert np.abs(lp - z["ref_logp"]).max() < 2e-5        # fp32 CPU vs the reference module's fp32 CPU
    if "tap_pool1_2" in z.files:                           # intermediate taps are stored for the first group of fixtures only
        assert np.abs(taps["pool1_2"].numpy()[:16] - z["tap_pool1_2"]).max() < 1e-5
    # calibrator apply restatement vs the fixture computed at generation time
    prob = torch.softmax(torch.from_numpy(z["ref_logp"]), 1).numpy()
    assert np.allclose(NT.dirichlet_apply(z["cal_weights"], prob), z["cal_prob"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag", INDEL_TAGS)
def test_unet_oracle_matches_reference(kat, tag):
    import os
    z = np.load(os.path.join(GOLD, "indel_%s.npz" % tag))
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    _, genome = kat
    names = list(genome)
    syms = [E.seq_to_symbols(genome[n]) for n in names]
    ch, st, sd = z["chrom"], z["start"], z["strand"]
    Rd = int(z["distal_radius"])
    oh = np.empty((len(st), 4, 2 * Rd), dtype=np.float32)
    for c in range(len(names)):
        m = ch == c
        oh[m] = E.onehot_windows(syms[c], st[m], sd[m], Rd, "indel")
    with torch.no_grad():
        o = NT.unet_small_forward(state, oh, [int(v) for v in z["down"]], bool(z["use_reverse"]), torch.float32).numpy()
    assert np.abs(o - z["ref_out"]).max() < 1e-4


def test_poisson_calibrate_properties():
    rng = np.random.default_rng(3)
    p = rng.dirichlet([50, 1, 1, 1], 100)
    q = NT.poisson_calibrate(p)
    lam = -np.log(p[:, 0])
    assert np.allclose(q[:, 0], 1 - lam)
    assert np.allclose(q[:, 1:].sum(1), lam)          # non-reference classes re-scaled to sum to lambda


def test_evaluation_oracle_matches_reference_goldens():
    """oracle/evaluation_np.py vs the outputs of the reference's own evaluation.py (tests/golden/eval_kat.npz)."""
    from oracle import evaluation_np as EN
    z = np.load(os.path.join(GOLD, "eval_kat.npz"))
    for tag, K, kmers in (("snv_f32", 4, [3, 5, 7]), ("snv_f64", 4, [3, 5, 7]), ("indel_f32", 8, [2, 4, 6])):
        flank, labels, prob = z[tag + ":flank"].astype(np.int64), z[tag + ":labels"].astype(np.int64), z[tag + ":prob"]
        f32 = prob.dtype == np.float32
        for k in kmers:
            assert np.allclose(EN.freq_kmer_comp_multi(flank, labels, prob, k, K, f32), z["%s:kmer%d" % (tag, k)], rtol=0, atol=1e-12, equal_nan=True)
        assert np.allclose(EN.calc_avg_prob(labels, prob, K, f32), z[tag + ":avg_prob"], rtol=0, atol=1e-15)
        if tag != "snv_f32":     # the float32 python loops are slow; one float32 case is covered by the indel fixture
            score, corr_list, n_regions = EN.regional_score(flank, labels, prob, len(prob), kmers, K, f32)
            assert n_regions == int(z[tag + ":regional_score"][1])
            assert abs(score - z[tag + ":regional_score"][0]) < 1e-9 * max(1.0, score)
            assert np.allclose(corr_list, z[tag + ":regional_corr_list"], rtol=0, atol=1e-10, equal_nan=True)
        names, chrom, start = z[tag + ":chrom_names"], z[tag + ":chrom"].astype(np.int64), z[tag + ":start"].astype(np.int64)
        order = np.lexsort((start, names[chrom]))
        for w in (100000, 500000):
            got = EN.corr_calc_sub(chrom[order], start[order], labels[order], prob[order], w, K, f32)
            assert np.allclose(got, z["%s:window%d" % (tag, w)], rtol=0, atol=1e-9, equal_nan=True)


def test_train_mode_oracle_matches_reference_goldens():
    """TRAIN mode (batch-statistic BatchNorm, dropout 0): output, CE(sum) loss, every parameter gradient and the updated
    running statistics of the oracle vs the unmodified reference run in float64 (tests/golden/train_kat.npz, written by
    oracle/make_golden_train.py).  The CUDA training paths are tested against this oracle's autograd."""
    z = np.load(os.path.join(GOLD, "train_kat.npz"))
    cases = [("snv_ex_ckpt6", "snv_ex_ckpt6.npz", None), ("indel_hs_ins", "indel_hs_ins.npz", True),
             ("indel_hs_del_start", "indel_hs_del_start.npz", True)]
    for tag, wfile, is_indel in cases:
        w = np.load(os.path.join(GOLD, wfile))
        state = {k[2:]: np.asarray(w[k]) for k in w.files if k.startswith("w:") and "num_batches" not in k}
        sd = {k: torch.tensor(v.astype(np.float64), requires_grad=("running" not in k)) for k, v in state.items()}
        x, y = z[tag + ":x"].astype(np.float64), z[tag + ":y"].astype(np.int64)
        rec = NT._BNStats()
        if is_indel:
            out = NT.unet_small_forward(sd, x, [int(v) for v in w["down"]], bool(w["use_reverse"]), torch.float64, train=True, rec=rec)
        else:
            out = NT.network2_forward(sd, z[tag + ":cat"].astype(np.int64), x, torch.float64, train=True, rec=rec)
        assert np.abs(out.detach().numpy() - z[tag + ":out"]).max() < 1e-9
        loss = NT.ce_sum(out, y)
        assert abs(float(loss.detach()) - float(z[tag + ":loss"])) < 1e-9 * max(1.0, abs(float(z[tag + ":loss"])))
        loss.backward()
        n_g = 0
        for k in z.files:
            if k.startswith(tag + ":g:"):
                name = k[len(tag) + 3:]
                g_ref = z[k]
                assert np.abs(sd[name].grad.numpy() - g_ref).max() <= 1e-7 * max(1e-6, np.abs(g_ref).max()), (tag, name)
                n_g += 1
        assert n_g > 100
        for bn, (mean, var_unb) in rec.stats.items():
            if (tag + ":rm:" + bn) not in z.files:
                continue
            em = 0.9 * state[bn + ".running_mean"].astype(np.float64) + 0.1 * mean.numpy()
            ev = 0.9 * state[bn + ".running_var"].astype(np.float64) + 0.1 * var_unb.numpy()
            assert np.abs(em - z[tag + ":rm:" + bn]).max() < 1e-9 * max(1, np.abs(em).max())
            assert np.abs(ev - z[tag + ":rv:" + bn]).max() < 1e-9 * max(1, np.abs(ev).max())
