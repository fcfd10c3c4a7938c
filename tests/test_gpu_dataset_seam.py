"""GPU: the dataset seam with the real encoders — the attributes the reference's train() / run_predict_pipline read
off the dataset equal the oracle's, and the reference-tuple route (DataLoader(ds,1) -> generate_data_batches -> model_predict_m
on tensors) gives exactly the predictions of the site-record fast path.  The build-container twin (tests/test_dataset_seam.py)
checks the same attributes against the reference's own CombinedDatasetNP and drives the reference's own loop."""
import numpy as np
import pytest
import torch

from conftest import load_snv_golden
from oracle import encode_np as E
from test_gpu_snv_forward import build_model

pytestmark = pytest.mark.gpu


def test_dataset_attributes_and_tuple_route(kat, cuda_genome):
    from torch.utils.data import DataLoader
    from mural_b200 import PackedSiteDataset, SiteTable, generate_data_batches, generate_site_batches, model_predict_m
    z, cfg, state = load_snv_golden("hs_AT")
    _, genome = kat
    names = list(genome)
    lab = z["start"] % 4
    t = SiteTable(names, z["chrom"], z["start"], z["start"] + 1, z["strand"], lab)
    ds = PackedSiteDataset(t, cuda_genome, 4000, cfg["local_radius"], cfg["local_order"], cfg["distal_radius"])
    # oracle local matrices in emission order
    ch, st, sd = z["chrom"][ds.perm], z["start"][ds.perm].astype(np.int64), z["strand"][ds.perm]
    for order, cols in ((1, ds.seq_cols), (cfg["local_order"], ds.cat_cols)):
        exp = np.empty((len(st), len(cols)), np.int64)
        for c in range(len(names)):
            m = ch == c
            exp[m] = E.kmer_windows(E.seq_to_symbols(genome[names[c]]), st[m], sd[m], cfg["local_radius"], order)
        assert np.array_equal(ds.local_codes(order), exp)
    assert ds.seq_cols[:2] == ["us7", "us6"] and ds.seq_cols[7] == "mid" and ds.cat_cols == ["cat%d" % (i + 1) for i in range(13)]
    assert ds.cat_dims == [int(v) + 1 for v in ds.local_codes(3).max(0)] and ds.cont_cols == []
    dl = ds.data_local
    assert list(dl.columns) == ds.seq_cols + ["mut_type"] and len(dl) == len(st)
    assert np.array_equal(dl["mut_type"].values, lab[ds.perm].astype(np.float64))
    assert dl.index.get_level_values(0).max() == len(ds) - 1
    # tuple route == fast route
    m = build_model(cfg, state, int(z["n_cat"]))
    fast, loss_fast = model_predict_m(m, generate_site_batches(ds, 2, 16), None, torch.device("cuda"), 4)
    crit = torch.nn.CrossEntropyLoss(reduction="sum")
    loader = generate_data_batches(DataLoader(ds.reference_view(), 1, shuffle=False), 2, 16, shuffle=False)
    slow, loss_slow = model_predict_m(m, loader, crit, torch.device("cuda"), 4)
    assert torch.equal(fast, slow)
    assert abs(loss_fast - loss_slow) <= 1e-4 * max(1.0, abs(loss_fast))


def test_mixed_focal_bases_exit_like_the_reference(kat, cuda_genome):
    """process_local_seq_snv (preprocessing.py:479-484): A/T and C/G focal bases mixed in one (segment, strand) batch -> exit."""
    from mural_b200 import PackedSiteDataset, SiteTable
    _, genome = kat
    seq = genome["chrB"].upper()
    a, c = seq.index("A", 500), seq.index("C", 500)
    t = SiteTable(list(genome), [1, 1], sorted([a, c]), [a + 1, c + 1], [0, 0], [0, 0])
    ds = PackedSiteDataset(t, cuda_genome, 300000, 7, 3, 200)
    with pytest.raises(SystemExit):
        ds.data_local


def test_label_range_checked():
    from mural_b200 import PackedSiteDataset, SiteTable

    class G:
        chrom_index = {"c": 0}
    t = SiteTable(["c"], [0, 0], [5, 9], [6, 10], [0, 0], [0, 200])
    with pytest.raises(ValueError):
        PackedSiteDataset(t, G(), 1000, 7, 3, 200)
