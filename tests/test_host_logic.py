"""CPU: host-side logic of mural_b200 (ordering, BED/FASTA ingest, batching, state_dict contract)."""
import gzip

import numpy as np
import pytest
import torch

from oracle import encode_np as E


def test_segment_order_matches_oracle_state_machine(kat):
    from mural_b200.data import segment_order
    z, _ = kat
    rng = np.random.default_rng(5)
    for central in (1, 13, 999, 5000, 300000):
        perm, sizes = segment_order(z["chrom"], z["start"], z["strand"], central)
        lit = E.bed_batches(z["chrom"], z["start"], z["strand"], central)
        assert np.array_equal(perm, np.concatenate([np.array(b[0]) for b in lit]))
        assert list(sizes) == [len(b[0]) for b in lit]
    for _ in range(20):                              # ragged / unsorted / repeated chromosomes
        n = int(rng.integers(1, 300))
        ch = rng.integers(0, 4, n); st = rng.integers(0, 20000, n); sd = rng.integers(0, 2, n)
        c = int(rng.integers(1, 5000))
        perm, sizes = segment_order(ch, st, sd, c)
        lit = E.bed_batches(ch, st, sd, c)
        assert np.array_equal(perm, np.concatenate([np.array(b[0]) for b in lit]))
    perm, sizes = segment_order([], [], [], 10)
    assert len(perm) == 0 and len(sizes) == 0


def test_segment_order_c_pass_equals_array_form():
    """mural_segment_order (one pass, stable partition per window run) == the stable sort by (block, window, strand)."""
    from mural_b200.data import segment_order, segment_order_np
    rng = np.random.default_rng(11)
    for trial in range(400):
        n = int(rng.integers(0, 400)); nchr = int(rng.integers(1, 5))
        ch = np.sort(rng.integers(0, nchr, n)) if trial % 3 else rng.integers(0, nchr, n)      # re-appearing chromosomes too
        st = rng.integers(0, 8000, n)
        if trial % 4:
            st = np.sort(st)
        sd = rng.integers(0, 2, n)
        c = int(rng.integers(1, 1200))
        a, b = segment_order(ch, st, sd, c), segment_order_np(ch, st, sd, c)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (trial, n, c)
    with pytest.raises(RuntimeError):
        segment_order([0], [5], [0], 0)


def test_site_records_c_pass_equals_array_form():
    """mural_pack_sites (gathers by the emission order + MURAL_META in one pass) == the numpy expressions it replaced."""
    from mural_b200.data import PackedSiteDataset, SiteTable, pack_meta

    class G:
        chrom_index = {"chrB": 2, "chrA": 0, "chrC": 1}
    rng = np.random.default_rng(3)
    gidx = np.array([2, 0, 1])
    for trial in range(60):
        n = int(rng.integers(0, 300))
        chrom = np.sort(rng.integers(0, 3, n)); start = rng.integers(0, 50000, n)
        strand = rng.integers(0, 2, n); label = rng.integers(0, 128, n)
        st = SiteTable(["chrB", "chrA", "chrC"], chrom[::1], start, start + 1, strand, label)
        ds = PackedSiteDataset(st, G(), int(rng.integers(1, 3000)), 10, 3, 1000)
        p = ds.perm
        assert ds.pos.dtype == np.int32 and np.array_equal(ds.pos, start[p])
        assert ds.strand.dtype == np.int8 and np.array_equal(ds.strand, strand[p])
        assert ds.label.dtype == np.int64 and np.array_equal(ds.label, label[p])
        assert np.array_equal(ds.chrom, gidx[chrom[p]])
        assert ds.meta.dtype == np.int32 and np.array_equal(ds.meta, pack_meta(ds.strand, ds.label, ds.chrom))
    bad = SiteTable(["chrA"], [0], [5], [6], [0], [200])
    with pytest.raises(ValueError):
        PackedSiteDataset(bad, G(), 100, 10, 3, 1000)


def test_bed_and_fasta_ingest(tmp_path):
    from mural_b200.data import SiteTable
    from mural_b200.genome import read_fasta
    bed = tmp_path / "a.bed.gz"
    with gzip.open(bed, "wt") as f:
        f.write("chr2\t10\t11\t.\t0\t+\nchr2\t15\t16\t.\t3\t-\nchr1\t7\t8\t.\t1\t+\n")
    t = SiteTable.from_bed(str(bed))
    assert t.chrom_names == ["chr2", "chr1"] and list(t.start) == [10, 15, 7]
    assert list(t.strand) == [0, 1, 0] and list(t.label) == [0, 3, 1]
    fa = tmp_path / "g.fa"
    fa.write_text(">chr1 desc\nACGT\nacgn\n>chr2\nTTTT\n")
    g = read_fasta(str(fa))
    assert g == {"chr1": b"ACGTacgn", "chr2": b"TTTT"}
    fa.write_text(">x\nAC\n>x\nGT\n")
    with pytest.raises(ValueError):
        read_fasta(str(fa))


class _FakeGenome:
    device = torch.device("cpu")
    chrom_index = {"chrA": 0, "chrB": 1, "chrC": 2}


def test_batches_preserve_reference_order(kat):
    """predict path: concatenated batches == bed_reader emission order (tail is carried to the next pool)."""
    from mural_b200.data import PackedSiteDataset, SiteTable, generate_site_batches
    z, genome = kat
    t = SiteTable(list(genome), z["chrom"], z["start"], z["start"] + 1, z["strand"], z["start"] % 4)
    ds = PackedSiteDataset(t, _FakeGenome(), 5000, 7, 3, 1000)
    for bs, pool in ((16, 1), (128, 10), (7, 3)):
        got = torch.cat([b.pos for b in generate_site_batches(ds, pool, bs, shuffle=False, device="cpu")]).numpy()
        assert np.array_equal(got, ds.pos)
        sizes = [len(b) for b in generate_site_batches(ds, pool, bs, shuffle=False, device="cpu")]
        assert all(s == bs for s in sizes[:-1]) and 0 < sizes[-1] <= bs
    # shuffled: a permutation of the same multiset
    got = torch.cat([b.pos for b in generate_site_batches(ds, 4, 32, shuffle=True, seed=1, device="cpu")]).numpy()
    assert np.array_equal(np.sort(got), np.sort(ds.pos))
    # meta packing round-trips
    assert np.array_equal(ds.meta & 1, ds.strand) and np.array_equal((ds.meta >> 1) & 0x7f, ds.label)
    assert np.array_equal(ds.meta >> 8, ds.chrom)


def test_state_dict_contract(manifest):
    """Same keys, shapes and dtypes as the reference Network2 (incl. '.layer.N.' aliases)."""
    from mural_b200 import model_choice
    cfg = {"local_radius": 7, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": 1000,
           "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": 3, "CNN_out_channels": 32, "distal_fc_dropout": .25,
           "n_class": 4, "model_no": 2}
    common = dict(emb_dims=[(65, 2)] * 13, n_cont=0, n_class=4, distal_order=1, in_channels=4)
    m = model_choice(2, cfg, common, "snv")
    ref = manifest["state_dict_keys"]["hs_AT"]
    sd = m.state_dict()
    assert [k for k, _, _ in ref] == list(sd.keys())
    for k, shape, dt in ref:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dt, k
    assert sum(p.numel() for p in m.parameters()) == 86904
    with pytest.raises(ValueError):
        model_choice(1, cfg, common, "snv")


def test_transfer_initialisation_mirrors_reference():
    """training.py:289-320: pretrained state loaded at another distal radius (all tensors are radius-independent); the two
    switches behave as in the reference, including its dead partial-freeze / fc re-init branches (AttributeError)."""
    import torch
    from mural_b200 import model_choice, weights_init
    from mural_b200.training import load_pretrained
    cfg = {"local_radius": 7, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": 1000,
           "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": 3, "CNN_out_channels": 32, "distal_fc_dropout": .25,
           "n_class": 4, "model_no": 2}
    common = dict(emb_dims=[(65, 2)] * 13, n_cont=0, n_class=4, distal_order=1, in_channels=4)
    torch.manual_seed(0)
    src = model_choice(2, cfg, common, "snv")
    src.apply(weights_init)
    state = {k: v.clone() for k, v in src.state_dict().items()}
    dst = model_choice(2, dict(cfg, distal_radius=200), common, "snv")
    dst.apply(weights_init)
    for p in dst.parameters():
        p.requires_grad = False
    load_pretrained(dst, state, train_all=True, init_fc_with_pretrained=True)
    assert all(p.requires_grad for p in dst.parameters())
    for k, v in dst.state_dict().items():
        assert torch.equal(v, state[k]), k
    with pytest.raises(AttributeError):
        load_pretrained(dst, state, train_all=False, init_fc_with_pretrained=True)
    with pytest.raises(AttributeError):
        load_pretrained(dst, state, train_all=True, init_fc_with_pretrained=False)
    with pytest.raises(RuntimeError):          # strict load: a state dict of another architecture is refused
        load_pretrained(dst, {k: v for k, v in state.items() if "conv3" not in k})
    with pytest.raises(SystemExit):
        load_pretrained(dst, state, train_all=False, model_type="indel")


def test_native_tsv_writer_equals_pandas(tmp_path):
    """mural_write_tsv (threaded C formatter) == DataFrame.sort_values(['chrom','start']).to_csv(sep='\\t',
    float_format='%.4g', index=False) byte for byte (run_predict.py:237-239), incl. ties, tiny / huge / exact values."""
    from mural_b200.predict import format_predictions, write_tsv
    rng = np.random.default_rng(0)
    n = 30000
    names = np.array(["chr10", "chr2", "chrX", "chr1"], dtype=object)[rng.integers(0, 4, n)]
    start = rng.integers(0, 50_000, n)            # many (chrom, start) ties: the sort must be stable
    end = start + 1
    strand = np.where(rng.integers(0, 2, n) == 0, "+", "-")
    mt = rng.integers(0, 4, n)
    prob = np.c_[rng.random(n), rng.random(n) * 1e-3, rng.random(n) * 1e-7, 10 ** rng.uniform(-12, 0, n)]
    prob[:5] = [[1, 0, 0.5, 1e-5], [0.99995, 1e-4, 123456.7, 1e-300], [0.1, 0.25, 1e-5, 9.9995e-5], [0, 0, 0, 0],
                [1e-4, 99999, 100000, 0.00012345]]
    a, b = tmp_path / "native.tsv", tmp_path / "pandas.tsv"
    write_tsv(a, names, start, end, strand, mt, prob, n_threads=3)
    format_predictions(names, start, end, strand, mt, prob).to_csv(b, sep="\t", float_format="%.4g", index=False)
    assert a.read_bytes() == b.read_bytes()
    write_tsv(a, names[:0], start[:0], end[:0], strand[:0], mt[:0], prob[:0])      # empty input: header only
    assert a.read_text() == "chrom\tstart\tend\tstrand\tmut_type\tprob0\tprob1\tprob2\tprob3\n"


def _bed_py(path):
    """The reader the C library replaced (kept here as the parity reference)."""
    import gzip
    opener = gzip.open if str(path).endswith(".gz") else open
    names, index, rows = [], {}, []
    with opener(path, "rt") as f:
        for line in f:
            if not line.strip() or line.startswith(("#", "track", "browser")):
                continue
            p = line.rstrip("\n").split("\t")
            if len(p) < 3:
                p = line.split()
            if p[0] not in index:
                index[p[0]] = len(names); names.append(p[0])
            rows.append((index[p[0]], int(p[1]), int(p[2]), 0 if (len(p) > 5 and p[5] == "+") else 1,
                         int(float(p[4])) if len(p) > 4 and p[4] not in (".", "") else 0))
    return names, np.array(rows, dtype=np.int64).reshape(-1, 5)


@pytest.mark.parametrize("gz", [False, True])
def test_native_bed_reader(tmp_path, gz):
    import gzip
    from mural_b200.data import SiteTable
    rng = np.random.default_rng(5)
    lines = ["# comment", "track name=x", "browser position chr1", "", "chr2\t10\t11\t.\t3\t+", "chr1\t5\t6\tn\t1.0\t-",
             "chr2 7 8 . 2 +", "chr10\t1\t2", "chr1\t9\t10\t.\t.\t+", "chr3\t4\t5\tx\t0\t*\textra\tcols", "chrX\t100\t101\t.\t\t+\r"]
    for _ in range(5000):
        lines.append("chr%d\t%d\t%d\t.\t%d\t%s" % (rng.integers(1, 6), rng.integers(0, 10**8), rng.integers(0, 10**8), rng.integers(0, 4), "+-"[rng.integers(0, 2)]))
    path = tmp_path / ("s.bed.gz" if gz else "s.bed")
    (gzip.open if gz else open)(path, "wt").write("\n".join(lines) + "\n")
    names, rows = _bed_py(path)
    t = SiteTable.from_bed(path)
    assert t.chrom_names == names and len(t) == len(rows)
    assert (t.chrom == rows[:, 0]).all() and (t.start == rows[:, 1]).all() and (t.end == rows[:, 2]).all()
    assert (t.strand == rows[:, 3]).all() and (t.label == rows[:, 4]).all()
    bad = tmp_path / "bad.bed"
    bad.write_text("chr1\tx\t2\n")
    with pytest.raises(ValueError):
        SiteTable.from_bed(bad)
    with pytest.raises(RuntimeError):
        SiteTable.from_bed(tmp_path / "missing.bed")


def test_native_fasta_reader(tmp_path):
    import ctypes as C
    import gzip
    from mural_b200 import _lib
    from mural_b200.genome import read_fasta
    rng = np.random.default_rng(6)
    recs = {"chrA desc": "ACGTNRYacgtn" * 50, "chrB": "".join("ACGT"[i] for i in rng.integers(0, 4, 12345)), "empty": "", "chrC\tx": "GATTACA"}
    txt = "; junk before the first header\n"
    for k, v in recs.items():
        txt += ">" + k + "\n" + "".join(v[i:i + 60] + ("  \r\n" if i % 120 else "\n") for i in range(0, len(v), 60))
    for gz in (False, True):
        path = tmp_path / ("g.fa.gz" if gz else "g.fa")
        (gzip.open if gz else open)(path, "wt", newline="").write(txt)
        exp = read_fasta(path)
        L = _lib.lib()
        f = C.c_void_p()
        _lib.check(L.mural_fasta_read(str(path).encode(), C.byref(f)))
        got = {L.mural_fasta_name(f, i).decode(): C.string_at(L.mural_fasta_seq(f, i), L.mural_fasta_len(f, i)) for i in range(L.mural_fasta_n(f))}
        L.mural_fasta_destroy(f)
        assert list(got) == list(exp) == ["chrA", "chrB", "empty", "chrC"]
        assert got == exp
    dup = tmp_path / "dup.fa"
    dup.write_text(">a\nAC\n>a\nGT\n")
    f = C.c_void_p()
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().mural_fasta_read(str(dup).encode(), C.byref(f)))


def test_lr_schedulers_match_torch():
    """StepLR / StepLR2 (per batch, with the min_lr -> restart_lr rule) and ReduceLROnPlateau (per epoch) vs torch on a dummy
    optimizer (training.py:364-371, 444-450)."""
    from mural_b200.training import auto_weight_decay, make_scheduler
    rng = np.random.default_rng(3)
    for kind in ("StepLR", "StepLR2"):
        cfg = {"lr_scheduler": kind, "learning_rate": 1e-3, "batch_size": 4096, "LR_gamma": 0.5, "min_lr": 1e-5, "restart_lr": 2e-4}
        train_size = 4096 * 40
        opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=cfg["learning_rate"])
        if kind == "StepLR":
            ref = torch.optim.lr_scheduler.StepLR(opt, step_size=(5000 * 128) // cfg["batch_size"], gamma=cfg["LR_gamma"])
        else:
            ref = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=(cfg["min_lr"] / cfg["restart_lr"]) ** (1 / (train_size // cfg["batch_size"])))
        mine = make_scheduler(cfg, train_size)
        for _ in range(1500):
            opt.step(); ref.step()
            if opt.param_groups[0]["lr"] < cfg["min_lr"]:
                for g in opt.param_groups:
                    g["lr"] = cfg["restart_lr"]
            lr = mine.step()
            assert abs(lr - opt.param_groups[0]["lr"]) <= 1e-12 + 1e-9 * lr, kind
    cfg = {"lr_scheduler": "ROP", "learning_rate": 1e-3, "batch_size": 128}
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
    ref = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.2, patience=1, threshold=0.0001, min_lr=1e-7)
    mine = make_scheduler(cfg, 1000)
    for loss in np.r_[np.linspace(5, 4, 6), 4 + 0.01 * rng.random(12), np.linspace(3.9, 3.0, 4), 3.0 * np.ones(10)]:
        ref.step(float(loss))
        assert abs(mine.step(float(loss)) - opt.param_groups[0]["lr"]) < 1e-15
    assert abs(auto_weight_decay(0.1, 128, 20, 10_000_000) - (1 - 0.1 ** (128 / (20 * 10_000_000)))) < 1e-18
    with pytest.raises(ValueError):
        auto_weight_decay(1.5, 128, 20, 1000)


def test_evaluator_host_math_on_numpy_tables():
    """The host half of mural_b200.evaluation (means from the integer tables, Pearson over observed groups, regional score)
    on tables built with numpy in the kernel's layout, against the reference's outputs (tests/golden/eval_kat.npz)."""
    import os
    import numpy as np
    from conftest import GOLD
    from mural_b200 import evaluation as EV
    z = np.load(os.path.join(GOLD, "eval_kat.npz"))
    tag, K = "snv_f64", 4
    flank, labels, prob = z[tag + ":flank"].astype(np.int64), z[tag + ":labels"].astype(np.int64), z[tag + ":prob"]
    n = len(labels)

    def table(k, region_size):
        d, mid = k // 2, flank.shape[1] // 2
        cols = [mid - j for j in range(d, 0, -1)] + [mid + j for j in range(1, d + 1)]
        g = np.zeros(n, np.int64)
        for c in cols:
            g = g * 5 + flank[:, c]
        R = n // region_size
        t = np.zeros((R, 5 ** (2 * d), 1 + 2 * K), np.int64)
        r = np.arange(n) // region_size
        ok = r < R
        np.add.at(t, (r[ok], g[ok], 0), 1)
        np.add.at(t, (r[ok], g[ok], 1 + labels[ok]), 1)
        for c in range(K):
            np.add.at(t, (r[ok], g[ok], 1 + K + c), np.rint(prob[ok, c] * EV.SCALE).astype(np.int64))
        return t
    for k in (3, 5, 7):
        assert np.allclose(EV._kmer_corr(table(k, n)[0], K, False), z["%s:kmer%d" % (tag, k)], rtol=0, atol=1e-7)
    region_size = n // 10
    score = sum(float(np.sum((1 - EV._kmer_corr_regions(table(k, region_size), K, False)) ** 2)) for k in (3, 5))
    assert abs(score - z[tag + ":regional_score"][0]) < 1e-6 * score
    # a region with a single observed group, or a constant column, gives NaN like Series.corr
    t = np.zeros((2, 25, 1 + 2 * K), np.int64)
    t[0, 3] = [5, 5, 0, 0, 0, 1, 1, 1, 1]
    t[1, 3] = [5, 5, 0, 0, 0, 1, 1, 1, 1]
    t[1, 4] = [7, 7, 0, 0, 0, 2, 2, 2, 2]
    assert np.isnan(EV._kmer_corr_regions(t, K, False)).all()


def test_native_bed_reader_multi_slice(tmp_path):
    """Files larger than a few MB are parsed in line-aligned slices by a thread pool: rows, chromosome numbering (order of
    first appearance in the FILE, with chromosomes that re-appear in later slices) and the line number of a malformed
    record must not depend on the slicing."""
    import numpy as np
    from mural_b200.data import SiteTable
    rng = np.random.default_rng(12)
    n = 600_000
    names = np.array(["chr2L", "chrX_random_scaffold_0001", "chr10", "chrM"])
    order = np.concatenate([np.zeros(200_000, int), np.full(150_000, 2), np.full(50_000, 0), np.full(100_000, 1), np.full(100_000, 3)])
    start = rng.integers(0, 10**8, n)
    strand = rng.integers(0, 2, n)
    label = rng.integers(0, 4, n)
    lines = ["%s\t%d\t%d\t.\t%d\t%s" % (names[c], s, s + 1, l, "+-"[d]) for c, s, l, d in zip(order, start, label, strand)]
    path = tmp_path / "big.bed"
    path.write_text("# header\n" + "\n".join(lines) + "\n")
    assert path.stat().st_size > 12 << 20
    t = SiteTable.from_bed(str(path))
    assert t.chrom_names == ["chr2L", "chr10", "chrX_random_scaffold_0001", "chrM"]
    remap = np.array([0, 2, 1, 3])
    assert np.array_equal(t.chrom, remap[order]) and np.array_equal(t.start, start) and np.array_equal(t.end, start + 1)
    assert np.array_equal(t.strand, strand) and np.array_equal(t.label, label)
    bad = 555_555
    lines[bad] = "chr10\tnot_a_number\t5\t.\t0\t+"
    path.write_text("# header\n" + "\n".join(lines) + "\n")
    with pytest.raises(ValueError, match="line %d" % (bad + 2)):
        SiteTable.from_bed(str(path))


def test_eval_data_from_reference_frames():
    """EvalData.from_frame on the reference's data_local headers (get_local_header, preprocessing.py:358-375): snv frames carry a
    `mid` column, indel frames do not; either way us_j / ds_j end up j columns left / right of the centre."""
    import numpy as np
    import pandas as pd
    from mural_b200.evaluation import EvalData
    rng = np.random.default_rng(0)
    n, R = 50, 4
    us, ds = ["us%d" % i for i in range(R, 0, -1)], ["ds%d" % i for i in range(1, R + 1)]
    for cols in (us + ["mid"] + ds, us + ds):
        df = pd.DataFrame(rng.integers(0, 5, (n, len(cols))), columns=cols)
        df["mut_type"] = rng.integers(0, 3, n)
        for i in range(3):
            df["prob%d" % i] = rng.random(n).astype(np.float32)
        ed = EvalData.from_frame(df, 3, device="cpu")
        flank = ed.flank.numpy()
        assert flank.shape == (n, 2 * R + 1) and ed.f32 and ed.prob.dtype.is_floating_point
        mid = R
        for j in range(1, R + 1):
            assert np.array_equal(flank[:, mid - j], df["us%d" % j].to_numpy()) and np.array_equal(flank[:, mid + j], df["ds%d" % j].to_numpy())
        assert np.array_equal(ed.labels_host(), df["mut_type"].to_numpy())


def test_tsv_float_formatter_equals_printf():
    """The TSV writer's own "%.4g" (one multiplication by a power of ten, printf only next to rounding ties) must be byte-identical
    to printf — which is what pandas' float_format applies per value (run_predict.py:236) — on probabilities of every magnitude,
    exact short decimals, rounding boundaries, negative values, zeros and non-finite values."""
    import ctypes as C
    from mural_b200 import _lib
    L = _lib.lib()
    buf = C.create_string_buffer(64)

    def f(v):
        L.mural_format_g4(float(v), buf)
        return buf.value.decode()
    special = [0.0, -0.0, 1.0, 2.0, 3.0, 0.5, 0.25, 0.1, 0.12345, 0.123449999, 0.99995, 0.99994999, 9999.5, 9999.4999, 99995.0, 1e-4,
               9.9995e-5, 1e-5, 1.2345e-5, 123456789.0, 1e19, 1e-19, 1e-20, 5e-324, 1e300, float("inf"), -float("inf"), -1.5e-7, 0.00012345,
               1234.5, 1234.4999999999, 0.1 + 0.2, 1 / 3, 2 / 3, 0.952381, 1e4, 9999.0, 1000.0, 0.001, 12.5, 100.0]
    for v in special:
        assert f(v) == "%.4g" % v, (v, f(v), "%.4g" % v)
    assert f(float("nan")).lstrip("-") == "nan"
    rng = np.random.default_rng(0)
    for xs in (rng.random(20000), rng.random(20000) * 1e-4, rng.random(20000) * 1e-8,
               10 ** rng.uniform(-22, 22, 40000) * rng.choice([-1, 1], 40000),
               np.arange(10000, 100000, 37)[:, None] * 10.0 ** np.array([-9, -8, -5, -4, -1, 0])[None, :]):
        for v in np.asarray(xs).ravel():
            assert f(v) == "%.4g" % v, (repr(float(v)), f(v), "%.4g" % v)
