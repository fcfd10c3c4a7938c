"""ORACLE — TEST INFRASTRUCTURE ONLY.  SURVEY 8f N4: Network2 with continuous features (n_cont > 0: first_bn_layer + the wider
first Linear of the local branch, model_snv.py:326-334,457-463) and the window means of bigWig tracks that feed them
(get_mean_bw_for_bed, preprocessing.py:725-750).  No shipped checkpoint has n_cont > 0, so the UNMODIFIED reference Network2
is instantiated through its own model_choice with n_cont = 2, randomly initialised (its own weights_init, seeded) with
non-trivial BatchNorm statistics, and run on KAT-genome sites with random cont_x; the oracle is pinned to it in the same pass.
The bigWig half is pinned on the reference's arithmetic with a stand-in for pyBigWig (a dense per-base array with NaN gaps):
get_mean_bw_for_bed itself runs unmodified.   Run in the build container only:   python -m oracle.make_golden_cont
"""
import json
import os
import sys
import types

import numpy as np
import torch

from oracle import network_t as NT
from oracle import ref_import as R
from oracle.make_golden import GOLD, clean_state, make_kat_genome, oracle_encode, pick_sites, to_regions


class FakeBigWig:
    """pyBigWig file stand-in: values(chrom, start, end, numpy=True) over dense per-base arrays (NaN = no data)."""
    tracks = {}

    def __init__(self, path):
        self.t = FakeBigWig.tracks[path]

    def chroms(self, c):
        return len(self.t[c])

    def values(self, c, a, b, numpy=True):
        return self.t[c][a:b].copy()


def main():
    assert R.available()
    pre, snv, indel, nnu = R.import_reference()
    genome = make_kat_genome()
    names = list(genome)
    rng = np.random.default_rng(77)
    cfg = {"local_radius": 7, "local_order": 3, "local_hidden1_size": 150, "local_hidden2_size": 75, "distal_radius": 300,
           "emb_dropout": .1, "local_dropout": .1, "CNN_kernel_size": 3, "CNN_out_channels": 32, "distal_fc_dropout": .25,
           "n_class": 4, "model_no": 2}
    n_cat, n_cont = 13, 2
    common = dict(emb_dims=[(65, 2)] * n_cat, n_cont=n_cont, n_class=4, distal_order=1, in_channels=4)   # without_bw_distal (training.py:257-260)
    torch.manual_seed(5)
    model = nnu.model_choice(2, cfg, common, "snv")
    model.apply(nnu.weights_init)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm1d) and mod.num_features > 0:
            mod.running_mean.normal_(0, .3); mod.running_var.uniform_(.5, 1.5)
            mod.weight.data.uniform_(.5, 1.5); mod.bias.data.normal_(0, .2)
    model.eval()
    ch, stt, sd_ = pick_sites(genome, rng, per_chrom=40)
    perm, sizes, cat, oh = oracle_encode(genome, ch, stt, sd_, 5000, 7, 3, 300, "snv")
    # ---- bigWig tracks: dense arrays with gaps; the reference's get_mean_bw_for_bed over a pyBigWig stand-in
    tracks = []
    for ti in range(n_cont):
        t = {}
        for nme in names:
            v = rng.gamma(2.0, 3.0, len(genome[nme])).astype(np.float32)
            for s in rng.integers(0, max(1, len(v) - 400), 6):
                v[s:s + int(rng.integers(1, 400))] = np.nan                  # uncovered stretches
            t[nme] = v
        tracks.append(t)
        FakeBigWig.tracks["track%d.bw" % ti] = t
    pre.pyBigWig = types.SimpleNamespace(open=FakeBigWig)
    radii = [50, 1000]
    bt = to_regions(names, ch, stt, sd_)                                       # FILE order (bed order)
    bw = pre.get_mean_bw_for_bed(["track0.bw", "track1.bw"], ["t0", "t1"], radii, bt, model_type="snv")
    cont_file = bw.values.astype(np.float64)                                   # [n, 2] in file order
    # the reference concatenates these rows positionally with the emission-ordered local frame (preprocessing.py:429-432)
    cont = cont_file.astype(np.float32)
    with torch.no_grad():
        ref_lp = model.forward((torch.from_numpy(cont), torch.from_numpy(cat)), torch.from_numpy(oh)).numpy()
        state = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items() if ".layer." not in k}
        o32 = NT.network2_forward(state, cat, oh, torch.float32, cont_x=cont).numpy()
    d32 = float(np.abs(o32 - ref_lp).max())
    assert d32 < 2e-5, d32
    out = {"cfg_json": np.array(json.dumps(cfg)), "n_cat": np.array(n_cat), "n_cont": np.array(n_cont), "chrom": ch[perm], "start": stt[perm],
           "strand": sd_[perm], "file_chrom": ch, "file_start": stt, "file_strand": sd_, "cont": cont, "cont_file64": cont_file,
           "bw_radii": np.array(radii), "ref_logp": ref_lp}
    for ti, t in enumerate(tracks):
        for nme in names:
            out["track%d:%s" % (ti, nme)] = t[nme]
    for k, v in state.items():
        out["w:" + k] = v
    np.savez_compressed(os.path.join(GOLD, "snv_cont_kat.npz"), **out)
    print("n_cont=2 Network2: reference vs oracle fp32 %.2e on %d sites; cont means in [%.3f, %.3f]" % (d32, len(perm), cont.min(), cont.max()))


if __name__ == "__main__":
    main()
