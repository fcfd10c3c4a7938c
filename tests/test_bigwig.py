"""CPU: bigWig reader + window means vs the UNMODIFIED reference's get_mean_bw_for_bed outputs recorded by
oracle/make_golden_cont.py (run over a dense stand-in for pyBigWig), through files written by tests/bigwig_writer.py in all
three section types, deflated and plain."""
import os

import numpy as np
import pytest

from bigwig_writer import write_bigwig
from conftest import GOLD


def _tracks(z, ti):
    names = [str(n) for n in np.load(os.path.join(GOLD, "encode_kat.npz"))["genome_names"]]
    return {n: z["track%d:%s" % (ti, n)] for n in names}


@pytest.mark.parametrize("compress", [True, False])
def test_window_means_equal_reference(tmp_path, compress):
    from mural_b200.bigwig import BigWig, mean_bw_for_sites
    from mural_b200.data import SiteTable
    z = np.load(os.path.join(GOLD, "snv_cont_kat.npz"))
    paths = []
    for ti in range(int(z["n_cont"])):
        p = str(tmp_path / ("track%d.bw" % ti))
        write_bigwig(p, _tracks(z, ti), compress=compress, items_per_section=300 + 211 * ti)
        paths.append(p)
    bw = BigWig(paths[0])
    t0 = _tracks(z, 0)
    assert bw.chroms == {n: len(v) for n, v in t0.items()}
    # dense read-back: every base through 1-bp windows
    for n, v in t0.items():
        got = bw.window_means(n, np.arange(len(v)), np.arange(len(v)) + 1)
        assert np.array_equal(got, np.nan_to_num(v).astype(np.float64)), n
    bw.close()
    names = list(t0)
    sites = SiteTable(names, z["file_chrom"], z["file_start"], z["file_start"] + 1, z["file_strand"], 0 * z["file_start"])
    got = mean_bw_for_sites(paths, [int(r) for r in z["bw_radii"]], sites)
    ref = z["cont_file64"]
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-6 * max(1.0, np.abs(ref).max())      # the reference averages float32 values with np.mean
    with pytest.raises(RuntimeError):
        BigWig(str(tmp_path / "missing.bw"))
    bad = tmp_path / "bad.bw"
    bad.write_bytes(b"\0" * 100)
    with pytest.raises(RuntimeError):
        BigWig(str(bad))


def test_dataset_carries_cont_columns_in_emission_order(tmp_path):
    from mural_b200.data import PackedSiteDataset, SiteTable, prepare_dataset_np
    z = np.load(os.path.join(GOLD, "snv_cont_kat.npz"))
    paths = []
    for ti in range(2):
        p = str(tmp_path / ("t%d.bw" % ti))
        write_bigwig(p, _tracks(z, ti))
        paths.append(p)
    names = list(_tracks(z, 0))

    class G:
        chrom_index = {n: i for i, n in enumerate(names)}
    sites = SiteTable(names, z["file_chrom"], z["file_start"], z["file_start"] + 1, z["file_strand"], 0 * z["file_start"])
    from mural_b200.bigwig import mean_bw_for_sites
    cont = mean_bw_for_sites(paths, [50, 1000], sites)
    ds = PackedSiteDataset(sites, G(), 5000, 7, 3, 300, cont_data=cont, cont_names=["t0", "t1"])
    assert ds.cont_cols == ["t0", "t1"] and ds.cont_X.shape == (len(sites), 2) and ds.cont_X.dtype == np.float32
    assert np.array_equal(ds.cont_X, cont[ds.perm].astype(np.float32))
