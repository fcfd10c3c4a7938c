"""ORACLE — TEST INFRASTRUCTURE ONLY.  BASELINE config 1: the bundled example BED
`examples/snv/data/validation.sorted.bed` (84 000 sites on chr2L, both strands, labels 0-3) predicted with the example
checkpoint `examples/snv/models/checkpoint_6` (R_l 7, 3-mers, R_d 200) on the SURVEY's synthetic chr2L: 23 100 000 bp iid
uniform ACGT from numpy default_rng(20221), then 'A' forced at every '+' site and 'T' at every '-' site of the BED (the real
dm6 chr2L is not shipped with the reference).  The sites travel as a compact fixture (the GPU box has no /root/reference);
the genome is regenerated from the seed by tests/config1.py.

Written: tests/golden/config1.npz = sites (start, strand, label in FILE order) + the log-probs of the UNMODIFIED reference
Network2 (fp32, CPU) in the reference's emission order + the oracle's calibrated probabilities, with the oracle pinned to the
reference in the same pass.   Run in the build container only:   python -m oracle.make_golden_config1
"""
import json
import os
import pickle

import numpy as np
import torch

from oracle import encode_np as E
from oracle import network_t as NT
from oracle import ref_import as R
from oracle.make_golden import GOLD, REF, cal_weights, clean_state

CHR2L_LEN = 23_100_000


def synth_chr2l(start, strand):
    rng = np.random.default_rng(20221)
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, CHR2L_LEN, dtype=np.uint8)].copy()
    seq[start[strand == 0]] = ord("A")
    seq[start[strand == 1]] = ord("T")
    return seq


def main():
    assert R.available()
    pre, snv, indel, nnu = R.import_reference()
    rows = [l.split("\t") for l in open(os.path.join(REF, "examples/snv/data/validation.sorted.bed")).read().splitlines()]
    start = np.array([int(r[1]) for r in rows], dtype=np.int64)
    strand = np.array([0 if r[5] == "+" else 1 for r in rows], dtype=np.int64)
    label = np.array([int(r[4]) for r in rows], dtype=np.int64)
    assert all(r[0] == "chr2L" for r in rows) and len(rows) == 84000
    seq = synth_chr2l(start, strand)
    rel = "examples/snv/models/checkpoint_6"
    cfg = pickle.load(open(os.path.join(REF, rel, "model.config.pkl"), "rb"))
    common = dict(emb_dims=cfg["emb_dims"], n_cont=0, n_class=cfg["n_class"], distal_order=1, in_channels=4)
    model = nnu.model_choice(cfg["model_no"], cfg, common, "snv")
    sd = torch.load(os.path.join(REF, rel, "model"), map_location="cpu")
    model.load_state_dict(sd)
    model.eval()
    state = clean_state(sd)
    central = int(cfg["segment_center"])
    chrom = np.zeros(len(start), np.int64)
    perm, sizes = E.order_sites(chrom, start, strand, central)
    sym = E._ASCII2SYM[seq]
    st_p, sd_p = start[perm], strand[perm]
    torch.set_num_threads(8)
    ref_lp = np.empty((len(perm), 4), np.float32)
    worst = 0.0
    with torch.no_grad():
        for b0 in range(0, len(perm), 2048):
            s, d = st_p[b0:b0 + 2048], sd_p[b0:b0 + 2048]
            cat = E.kmer_windows(sym, s, d, cfg["local_radius"], cfg["local_order"])
            oh = E.onehot_windows(sym, s, d, cfg["distal_radius"])
            ref_lp[b0:b0 + 2048] = model.forward((torch.zeros(len(s), 1), torch.from_numpy(cat)), torch.from_numpy(oh)).numpy()
            if b0 % (2048 * 8) == 0:      # the oracle on a sample of the batches
                o32 = NT.network2_forward(state, cat, oh, torch.float32).numpy()
                worst = max(worst, float(np.abs(o32 - ref_lp[b0:b0 + 2048]).max()))
    assert worst < 2e-5, worst
    # the reference's own encoders on the first segment batch (bit-exact pin of the oracle's windows on this genome)
    calw = cal_weights(os.path.join(REF, rel, "model.fdiri_cal.pkl"))
    prob = torch.softmax(torch.from_numpy(ref_lp), 1).numpy()
    out = {"start": start.astype(np.int32), "strand": strand.astype(np.int8), "label": label.astype(np.int8), "perm": perm.astype(np.int32),
           "batch_sizes": np.asarray(sizes, np.int32), "ref_logp": ref_lp, "cal_weights": calw, "cal_prob": NT.dirichlet_apply(calw, prob).astype(np.float32),
           "chr2l_len": np.array(CHR2L_LEN), "seed": np.array(20221),
           "cfg_json": np.array(json.dumps({k: (v if not isinstance(v, list) else None) for k, v in cfg.items() if k != "emb_dims"}, default=float))}
    np.savez_compressed(os.path.join(GOLD, "config1.npz"), **out)
    print("config 1: %d sites, %d segment batches, oracle vs reference %.2e, label counts %s, strands +%d -%d, mean p0 %.4f" %
          (len(perm), len(sizes), worst, np.bincount(label).tolist(), int((strand == 0).sum()), int((strand == 1).sum()), prob[:, 0].mean()))


if __name__ == "__main__":
    main()
