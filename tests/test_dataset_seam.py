"""CPU (build container): the dataset seam of SURVEY 8b, exercised by the REFERENCE's own consumers.

The reference's `prepare_dataset_np` product (CombinedDatasetNP) and this repo's PackedSiteDataset are built over the same
BED-like sites and genome; every attribute the reference's train() / run_predict_pipline read off the dataset
(training.py:166-168,250-255; run_predict.py:150,199-200,226-236) must be equal, and the reference's own
`DataLoader(ds, 1)` -> `generate_data_batches` -> `model_predict_m` must produce the same predictions from either dataset.
Without a GPU the window encoders behind the dataset are replaced by the oracle (test infrastructure); the GPU twin of this
test (tests/test_gpu_dataset_seam.py) runs the real encoders.  Skipped where /root/reference does not exist.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import encode_np as E
from oracle import ref_import as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present (GPU box)")


class OracleGenome:
    """Stands in for PackedGenome on a GPU-less host: same methods, windows from the numpy oracle."""
    device = torch.device("cpu")

    def __init__(self, genome):
        self.names = list(genome)
        self.chrom_index = {n: i for i, n in enumerate(self.names)}
        self.sym = [E.seq_to_symbols(genome[n]) for n in self.names]

    def _per_chrom(self, fn, pos, meta, width, dtype, extra):
        pos, meta = np.asarray(pos), np.asarray(meta)
        out = np.empty((len(pos),) + width, dtype=dtype)
        ch, sd = meta >> 8, meta & 1
        for c in np.unique(ch):
            m = ch == c
            out[m] = fn(self.sym[c], pos[m].astype(np.int64), sd[m].astype(np.int64), *extra)
        return torch.from_numpy(out)

    def encode_local(self, pos, meta, radius, order, model_type="snv"):
        return self._per_chrom(E.kmer_windows, pos, meta, (E.window_length(radius, model_type) - (order - 1),), np.int64,
                               (radius, order, model_type))

    def encode_onehot(self, pos, meta, radius, model_type="snv"):
        return self._per_chrom(E.onehot_windows, pos, meta, (4, E.window_length(radius, model_type)), np.float32, (radius, model_type))


def _sites(kat, n_keep=400):
    z, genome = kat
    keep = np.sort(np.random.default_rng(3).choice(len(z["start"]), n_keep, replace=False))
    ch, st, sd = z["chrom"][keep], z["start"][keep].astype(np.int64), z["strand"][keep]
    lab = (st % 4).astype(np.int64)
    return genome, ch, st, sd, lab


@pytest.mark.parametrize("central,order", [(5000, 3), (997, 1)])
def test_dataset_attributes_equal_reference(kat, central, order):
    from mural_b200.data import PackedSiteDataset, SiteTable
    pre, _, _, _ = R.import_reference()
    from oracle.make_golden import to_regions
    genome, ch, st, sd, lab = _sites(kat)
    names = list(genome)
    bt = to_regions(names, ch, st, sd, lab)
    recs = {k: R.SeqRec(v) for k, v in genome.items()}
    R_l, R_d = 7, 150
    data_local, seq_cols, cat_feats, out_col = pre.prepare_local_data(bt, recs, [], [], [], central, R_l, order, True, model_type="snv")
    ref = pre.CombinedDatasetNP(data=data_local, seq_cols=seq_cols, cat_cols=cat_feats, output_col=out_col, ref_genome=recs, bed_regions=bt,
                                central_radius=central, distal_radius=R_d, n_channels=4, bw_files=[], seq_only=True,
                                without_bw_distal=False, model_type="snv")
    ref.get_distal_encoding_infomation()
    mine = PackedSiteDataset(SiteTable(names, ch, st, st + 1, sd, lab), OracleGenome(genome), central, R_l, order, R_d)
    # --- attributes read by the reference's train() and run_predict_pipline
    assert mine.cat_cols == ref.cat_cols and mine.cont_cols == ref.cont_cols and mine.seq_cols == seq_cols
    assert mine.cat_dims == [int(v) for v in ref.cat_dims]
    assert len(mine) == len(ref) and mine.distal_info
    a, b = mine.data_local, ref.data_local
    assert list(a.columns) == list(b.columns) and a.index.equals(b.index)
    assert np.array_equal(a.values, b.values) and [str(t) for t in a.dtypes] == [str(t) for t in b.dtypes]
    assert mine.y.index.equals(ref.y.index) and np.array_equal(mine.y.values, ref.y.values) and mine.y.dtype == ref.y.dtype
    assert np.array_equal(mine.get_labels(), ref.get_labels())
    info = pre.get_position_info(bt, central)
    nm, s0, e0, sd0 = mine.position_info()
    assert list(nm) == list(info["chrom"]) and np.array_equal(s0, info["start"]) and list(sd0) == list(info["stand"])
    # --- item contract
    view = mine.reference_view()
    for i in (0, len(ref) // 2, len(ref) - 1):
        for u, v in zip(view[i], ref[i]):
            u, v = np.asarray(u), np.asarray(v)
            assert u.shape == v.shape and u.dtype == v.dtype and np.array_equal(u, v)


def test_reference_loop_consumes_the_dataset(kat):
    """The reference's own DataLoader -> generate_data_batches -> model_predict_m over this repo's dataset view gives the
    predictions it gives over its own dataset; this repo's generate_data_batches yields the same batches."""
    from torch.utils.data import DataLoader
    from mural_b200.data import PackedSiteDataset, SiteTable, generate_data_batches
    pre, snv, _, nnu = R.import_reference()
    from oracle.make_golden import to_regions
    genome, ch, st, sd, lab = _sites(kat, 300)
    names = list(genome)
    z = np.load(os.path.join(GOLD, "snv_ex_ckpt6.npz"))
    cfg = json.loads(str(z["cfg_json"]))
    R_l, order, R_d, central = cfg["local_radius"], cfg["local_order"], cfg["distal_radius"], 3000
    bt = to_regions(names, ch, st, sd, lab)
    recs = {k: R.SeqRec(v) for k, v in genome.items()}
    data_local, seq_cols, cat_feats, out_col = pre.prepare_local_data(bt, recs, [], [], [], central, R_l, order, True, model_type="snv")
    ref = pre.CombinedDatasetNP(data=data_local, seq_cols=seq_cols, cat_cols=cat_feats, output_col=out_col, ref_genome=recs, bed_regions=bt,
                                central_radius=central, distal_radius=R_d, n_channels=4, bw_files=[], seq_only=True,
                                without_bw_distal=False, model_type="snv")
    ref.get_distal_encoding_infomation()
    mine = PackedSiteDataset(SiteTable(names, ch, st, st + 1, sd, lab), OracleGenome(genome), central, R_l, order, R_d).reference_view()
    emb_dims = [(x, min(16, int(x ** 0.25))) for x in mine.cat_dims]                     # training.py:250-255
    common = dict(emb_dims=emb_dims, n_cont=len(mine.cont_cols), n_class=4, distal_order=1, in_channels=4)
    torch.manual_seed(0)
    model = nnu.model_choice(2, cfg, common, "snv")
    model.apply(nnu.weights_init)
    crit = torch.nn.CrossEntropyLoss(reduction="sum")
    outs = []
    for ds in (ref, mine):
        loader = pre.generate_data_batches(DataLoader(ds, 1, shuffle=False), 2, 16, shuffle=False)
        pred, loss = nnu.model_predict_m(model, loader, crit, torch.device("cpu"), 4, distal=True, model_type="snv")
        outs.append((pred.numpy(), loss))
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
    assert outs[0][0].shape == (len(st), 4)
    for pool, bs in ((2, 16), (3, 7), (100, 64)):
        a = list(pre.generate_data_batches(DataLoader(ref, 1, shuffle=False), pool, bs, shuffle=False))
        b = list(generate_data_batches(DataLoader(mine, 1, shuffle=False), pool, bs, shuffle=False))
        assert len(a) == len(b)
        for x, y in zip(a, b):
            for u, v in zip(x, y):
                assert u.shape == v.shape and u.dtype == v.dtype and torch.equal(u, v)
