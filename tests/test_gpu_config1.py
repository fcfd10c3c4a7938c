"""GPU: BASELINE config 1 — the reference's bundled example BED (84 000 sites on chr2L, both strands) predicted with
the example checkpoint on the SURVEY's synthetic chr2L (seeded; the '+' / '-' sites forced to A / T), end to end through
run_predict (FASTA + BED + checkpoint files -> calibrated probabilities) in the fp32-equivalent and the default auto mode, vs
the UNMODIFIED reference's log-probs recorded by oracle/make_golden_config1.py.  Gates (BASELINE north_star): 1e-3 / 5e-3 on
probabilities; sample order == the reference's emission order."""
import os
import pickle

import numpy as np
import pytest
import torch

from conftest import GOLD, load_snv_golden

pytestmark = pytest.mark.gpu


def _synth_chr2l(z):
    rng = np.random.default_rng(int(z["seed"]))
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, int(z["chr2l_len"]), dtype=np.uint8)].copy()
    seq[z["start"][z["strand"] == 0]] = ord("A")
    seq[z["start"][z["strand"] == 1]] = ord("T")
    return seq


def test_config1_example_bed_end_to_end(tmp_path):
    from mural_b200 import PackedGenome, PackedSiteDataset, SiteTable
    from mural_b200.predict import run_predict
    from test_gpu_predict_pipeline import _write_inputs
    from test_gpu_snv_forward import build_model
    z = np.load(os.path.join(GOLD, "config1.npz"))
    zc, cfg, state = load_snv_golden("ex_ckpt6")                    # the same checkpoint (weights + calibrator)
    seq = _synth_chr2l(z)
    fa, bed = tmp_path / "chr2L.fa", tmp_path / "validation.sorted.bed"
    with open(fa, "wb") as f:
        f.write(b">chr2L\n")
        body = seq[: len(seq) // 100 * 100].reshape(-1, 100)
        f.write(np.concatenate([body, np.full((len(body), 1), 10, np.uint8)], 1).tobytes())
        f.write(seq[len(seq) // 100 * 100:].tobytes() + b"\n")
    with open(bed, "w") as f:
        for s, d, l in zip(z["start"], z["strand"], z["label"]):
            f.write("chr2L\t%d\t%d\t.\t%d\t%s\n" % (s, s + 1, l, "+-"[d]))
    m = build_model(cfg, state, int(zc["n_cat"]))
    torch.save({k: v.cpu() for k, v in m.state_dict().items()}, tmp_path / "model")
    pickle.dump(dict(cfg, emb_dims=[(65, 2)] * int(zc["n_cat"]), segment_center=300000), open(tmp_path / "model.config.pkl", "wb"))
    # emission order == the reference's (bed_reader over 300 kb segments)
    genome = PackedGenome.from_fasta(str(fa))
    ds = PackedSiteDataset(SiteTable.from_bed(str(bed)), genome, 300000, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"])
    assert np.array_equal(ds.perm, z["perm"]) and np.array_equal(ds.batch_sizes, z["batch_sizes"])
    ref_prob = torch.softmax(torch.from_numpy(z["ref_logp"]), 1).numpy()
    inv = np.argsort(z["perm"])                                       # BED is (chrom, start)-sorted: output row i == file row i
    for mode, tol in (("fp32", 1e-3), ("auto", 5e-3)):
        df = run_predict(str(bed), str(fa), str(tmp_path / "model"), str(tmp_path / "model.config.pkl"), "", str(tmp_path / ("pred_%s.tsv" % mode)),
                         compute_mode=mode, genome=genome)
        assert len(df) == 84000 and np.array_equal(df["start"].values, z["start"]) and np.array_equal(df["mut_type"].values, z["label"])
        got = df[["prob0", "prob1", "prob2", "prob3"]].values
        d = np.abs(got - ref_prob[inv]).max()
        print("config 1, mode %s: max |dp| vs the reference = %.2e (gate %.0e); mean p0 %.4f" % (mode, d, tol, got[:, 0].mean()))
        assert d <= tol, (mode, d)
        assert os.path.getsize(tmp_path / ("pred_%s.tsv" % mode)) > 84000 * 30
