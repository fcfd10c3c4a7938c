"""GPU: BED + FASTA + checkpoint files -> calibrated TSV, vs the oracle pipeline (run_predict.py:214-239 restated)."""
import pickle
import sys
import types

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import load_snv_golden
from oracle import encode_np as E
from oracle import network_t as NT

pytestmark = pytest.mark.gpu


def _write_inputs(tmp_path, genome, z, cfg, state):
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as f:
        for n, s in genome.items():
            f.write(">%s test\n" % n)
            for i in range(0, len(s), 60):
                f.write(s[i:i + 60] + "\n")
    names = list(genome)
    order = np.lexsort((z["start"], z["chrom"]))            # BED sorted by chrom, start
    bed = tmp_path / "sites.bed"
    with open(bed, "w") as f:
        for i in order:
            f.write("%s\t%d\t%d\t.\t%d\t%s\n" % (names[z["chrom"][i]], z["start"][i], z["start"][i] + 1, z["start"][i] % 4, "+-"[z["strand"][i]]))
    from test_gpu_snv_forward import build_model
    m = build_model(cfg, state, int(z["n_cat"]))
    torch.save({k: v.cpu() for k, v in m.state_dict().items()}, tmp_path / "model")
    full_cfg = dict(cfg); full_cfg["emb_dims"] = [(65, 2)] * int(z["n_cat"]); full_cfg["segment_center"] = 5000
    pickle.dump(full_cfg, open(tmp_path / "model.config.pkl", "wb"))
    # a pickle shaped like the reference's model.fdiri_cal.pkl (FullDirichletCalibrator -> calibrator_.weights_)
    mod = types.ModuleType("dirichletcal.calib.fulldirichlet")
    FullDirichletCalibrator = type("FullDirichletCalibrator", (), {"__module__": "dirichletcal.calib.fulldirichlet"})
    MultinomialRegression = type("MultinomialRegression", (), {"__module__": "dirichletcal.calib.fulldirichlet"})
    FullDirichletCalibrator.__qualname__ = "FullDirichletCalibrator"; MultinomialRegression.__qualname__ = "MultinomialRegression"
    mod.FullDirichletCalibrator, mod.MultinomialRegression = FullDirichletCalibrator, MultinomialRegression
    for k in ("dirichletcal", "dirichletcal.calib"):
        sys.modules.setdefault(k, types.ModuleType(k))
    sys.modules["dirichletcal.calib.fulldirichlet"] = mod
    cal = FullDirichletCalibrator(); cal.calibrator_ = MultinomialRegression(); cal.calibrator_.weights_ = z["cal_weights"]
    pickle.dump(cal, open(tmp_path / "model.fdiri_cal.pkl", "wb"))
    for k in ("dirichletcal.calib.fulldirichlet", "dirichletcal.calib", "dirichletcal"):
        sys.modules.pop(k, None)
    return fa, bed, order


@pytest.mark.parametrize("poisson", [False, True])
def test_tsv_matches_oracle_pipeline(tmp_path, kat, poisson):
    from mural_b200.predict import run_predict
    z, cfg, state = load_snv_golden("hs_AT")
    _, genome = kat
    fa, bed, order = _write_inputs(tmp_path, genome, z, cfg, state)
    out = tmp_path / "pred.tsv"
    df = run_predict(str(bed), str(fa), str(tmp_path / "model"), str(tmp_path / "model.config.pkl"),
                     str(tmp_path / "model.fdiri_cal.pkl"), str(out), poisson_calib=poisson, compute_mode="fp32")
    # ---- oracle: reference order (bed_reader), network, softmax, calibrator, poisson, sort, %.4g
    names = list(genome)
    ch, st, sd = z["chrom"][order], z["start"][order], z["strand"][order]
    perm, _ = E.order_sites(ch, st, sd, 5000)
    ch, st, sd = ch[perm], st[perm], sd[perm]
    cat = np.empty((len(st), int(z["n_cat"])), np.int64); oh = np.empty((len(st), 4, 2 * cfg["distal_radius"] + 1), np.float32)
    for c in range(len(names)):
        msk = ch == c
        sym = E.seq_to_symbols(genome[names[c]])
        cat[msk] = E.kmer_windows(sym, st[msk], sd[msk], cfg["local_radius"], cfg["local_order"])
        oh[msk] = E.onehot_windows(sym, st[msk], sd[msk], cfg["distal_radius"])
    with torch.no_grad():
        lp = NT.network2_forward(state, cat, oh, torch.float32)
    prob = NT.dirichlet_apply(z["cal_weights"], torch.softmax(lp, 1).numpy())
    if poisson:
        prob = NT.poisson_calibrate(prob)
    exp = pd.DataFrame({"chrom": np.array(names, dtype=object)[ch], "start": st, "end": st + 1, "strand": np.where(sd == 0, "+", "-"),
                        "mut_type": (st % 4).astype(np.float64)})
    for i in range(4):
        exp["prob%d" % i] = prob[:, i]
    exp.sort_values(["chrom", "start"], inplace=True); exp.reset_index(drop=True, inplace=True)
    assert list(df.columns) == list(exp.columns)
    for c in ("chrom", "start", "end", "strand", "mut_type"):
        assert (df[c].values == exp[c].values).all(), c
    got_p, exp_p = df[["prob%d" % i for i in range(4)]].values, exp[["prob%d" % i for i in range(4)]].values
    assert np.abs(got_p - exp_p).max() <= 1e-3                  # fp32-equivalent gate on calibrated probabilities
    # file: same header / non-float columns; floats agree at the printed precision up to the gate
    exp_file = tmp_path / "exp.tsv"
    exp.to_csv(exp_file, sep="\t", float_format="%.4g", index=False)
    a, b = open(out).read().splitlines(), open(exp_file).read().splitlines()
    assert a[0] == b[0] and len(a) == len(b)
    same = sum(x == y for x, y in zip(a, b))
    assert same >= 0.97 * len(a), "only %d of %d lines are textually identical" % (same, len(a))
    for x, y in zip(a[1:], b[1:]):
        fx, fy = x.split("\t"), y.split("\t")
        assert fx[:5] == fy[:5]
        assert np.abs(np.array(fx[5:], float) - np.array(fy[5:], float)).max() <= 1e-3


def test_indel_tsv_matches_oracle_pipeline(tmp_path, kat):
    """MuRaL-indel through the same pipeline (run_predict.py with model_type='indel'): UNet_Small outputs -> softmax ->
    Poisson calibration (always on for indel, run_predict.py:224) -> sorted %.4g TSV."""
    import os
    from conftest import GOLD
    from mural_b200 import model_choice
    from mural_b200.predict import run_predict
    z = np.load(os.path.join(GOLD, "indel_ex_indel9.npz"))
    _, genome = kat
    names = list(genome)
    state = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    cfg = {"CNN_out_channels": int(state["uplblocks.0.0.weight"].shape[0]), "CNN_kernel_size": int(state["uplblocks.0.0.weight"].shape[2]),
           "down_list": [int(v) for v in z["down"]], "use_reverse": bool(z["use_reverse"]), "n_class": int(state["out_fc.2.weight"].shape[0]),
           "model_no": 0, "local_radius": 5, "local_order": 3, "distal_radius": int(z["distal_radius"]), "segment_center": 5000}
    m = model_choice(0, cfg, {"n_class": cfg["n_class"]}, "indel")
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in state.items()}, strict=True)
    torch.save(m.state_dict(), tmp_path / "model")
    pickle.dump(cfg, open(tmp_path / "model.config.pkl", "wb"))
    fa = tmp_path / "ref.fa"
    with open(fa, "w") as f:
        for n, s in genome.items():
            f.write(">%s\n%s\n" % (n, s))
    order = np.lexsort((z["start"], z["chrom"]))
    bed = tmp_path / "sites.bed"
    with open(bed, "w") as f:
        for i in order:
            f.write("%s\t%d\t%d\t.\t%d\t%s\n" % (names[z["chrom"][i]], z["start"][i], z["start"][i] + 1, z["start"][i] % 8, "+-"[z["strand"][i]]))
    out = tmp_path / "pred.tsv"
    df = run_predict(str(bed), str(fa), str(tmp_path / "model"), str(tmp_path / "model.config.pkl"), "", str(out), model_type="indel")
    ch, st, sd = z["chrom"][order], z["start"][order], z["strand"][order]
    perm, _ = E.order_sites(ch, st, sd, 5000)
    ch, st, sd = ch[perm], st[perm], sd[perm]
    oh = np.empty((len(st), 4, 2 * cfg["distal_radius"]), np.float32)
    for c in range(len(names)):
        msk = ch == c
        if msk.any():
            oh[msk] = E.onehot_windows(E.seq_to_symbols(genome[names[c]]), st[msk], sd[msk], cfg["distal_radius"], "indel")
    with torch.no_grad():
        y = NT.unet_small_forward(state, oh, cfg["down_list"], cfg["use_reverse"], torch.float32)
    prob = NT.poisson_calibrate(torch.softmax(y, 1).numpy().astype(np.float64))
    exp = pd.DataFrame({"chrom": np.array(names, dtype=object)[ch], "start": st, "end": st + 1, "strand": np.where(sd == 0, "+", "-"),
                        "mut_type": (st % 8).astype(np.float64)})
    for i in range(cfg["n_class"]):
        exp["prob%d" % i] = prob[:, i]
    exp.sort_values(["chrom", "start"], inplace=True); exp.reset_index(drop=True, inplace=True)
    cols = ["prob%d" % i for i in range(cfg["n_class"])]
    assert (df[["chrom", "start", "strand"]].values == exp[["chrom", "start", "strand"]].values).all()
    assert np.abs(df[cols].values - exp[cols].values).max() <= 1e-3
    got = pd.read_csv(out, sep="\t")
    assert list(got.columns) == list(exp.columns) and len(got) == len(exp)
    assert np.abs(got[cols].values - exp[cols].values).max() <= 1e-3 + 5e-4 * np.abs(exp[cols].values).max()


def test_auto_mode_routes_exception_windows_to_fp32(kat, cuda_genome):
    """compute_mode='auto' (MURAL_MODE_AUTO, inside the library): bf16 everywhere, the fp32-equivalent path for sites whose
    window has N / IUPAC symbols or overhangs the chromosome.  The device-built site list equals a brute-force scan of the
    windows; masked rows equal the fp32 path bit for bit, the others the bf16 path."""
    from mural_b200 import PackedSiteDataset, SiteTable
    from mural_b200.predict import predict_sites
    from test_gpu_snv_forward import build_model
    z, cfg, state = load_snv_golden("ex_ckpt6")
    _, genome = kat
    names = list(genome)
    rng = np.random.default_rng(31)
    n = 4000
    ch = rng.integers(0, len(names), n)
    st = np.array([rng.integers(0, len(genome[names[c]])) for c in ch])
    order = np.lexsort((st, ch)); ch, st = ch[order], st[order]
    sd = rng.integers(0, 2, n)
    sites = SiteTable(names, ch, st, st + 1, sd, st % 4)
    R = cfg["distal_radius"]
    ds = PackedSiteDataset(sites, cuda_genome, 5000, cfg["local_radius"], cfg["local_order"], R)
    mask = cuda_genome.windows_with_exceptions(ds.chrom, ds.pos, R)
    brute = np.array([(p - R < 0) or (p + R + 1 > len(genome[names[c]])) or any(b not in "ACGTacgt" for b in genome[names[c]][max(0, p - R):p + R + 1])
                      for c, p in zip(ds.chrom, ds.pos)])
    assert (mask == brute).all() and 0 < mask.sum() < n
    m = build_model(cfg, state, int(z["n_cat"]))
    out = {}
    from mural_b200 import _lib
    for mode in ("fp32", "bf16", "auto"):
        m.compute_mode = mode
        out[mode] = predict_sites(m, ds, 0, n).cpu().numpy()
    # MURAL_MODE_AUTO builds the same site list on the device (k_exception_sites): per chromosome call, so count the last one
    last_chrom = ds.chrom == ds.chrom[-1]
    assert int(_lib.lib().mural_snv_last_auto_sites(m._h)) == int(mask[last_chrom].sum())
    assert np.array_equal(out["auto"][mask], out["fp32"][mask])
    assert np.array_equal(out["auto"][~mask], out["bf16"][~mask])
    # the reference's tensor signature goes the same way (windows scanned for non-ACGT columns)
    import torch
    sel = np.r_[np.flatnonzero(mask)[:40], np.flatnonzero(~mask)[:40]]
    pos = torch.from_numpy(ds.pos[sel]).cuda(); meta = torch.from_numpy(ds.meta[sel]).cuda()
    cat = cuda_genome.encode_local(pos, meta, cfg["local_radius"], cfg["local_order"])
    oh = cuda_genome.encode_onehot(pos, meta, R)
    with torch.no_grad():
        t = m.forward((None, cat), oh).cpu().numpy()
    assert np.array_equal(t, out["auto"][sel])
