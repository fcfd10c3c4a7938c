"""GPU: validation metrics (csrc/metrics.cu through mural_b200.evaluation) vs the reference's own outputs (tests/golden/eval_kat.npz,
written by oracle/make_golden_eval.py from MuRaL/evaluation/evaluation.py) and vs the numpy oracle on larger inputs.

Tolerances: float64 probabilities 1e-7 on correlations (fixed-point sums, 2^-37 per site).  With float32 probabilities the
reference itself accumulates in float32 — pandas' Kahan group means (one float32 ulp: 2e-5 on k-mer correlations) and, in
corr_calc_sub, a plain sequential float32 sum over the ~1 300 sites of a 500 kb window, whose rounding noise moves the window
correlations by up to 3.6e-5 on these fixtures (exact float64 sums vs the reference's output, oracle/evaluation_np.py): 2e-4."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLD
from oracle import evaluation_np as EN

pytestmark = pytest.mark.gpu
TAGS = [("snv_f32", 4, [3, 5, 7]), ("snv_f64", 4, [3, 5, 7]), ("indel_f32", 8, [2, 4, 6])]


def _frames(z, tag, n_class):
    import pandas as pd
    flank = z[tag + ":flank"].astype(np.int64)
    R = flank.shape[1] // 2
    cols = ["us%d" % i for i in range(R, 0, -1)] + ["mid"] + ["ds%d" % i for i in range(1, R + 1)]
    data_local = pd.DataFrame(flank, columns=cols)
    data_local["mut_type"] = z[tag + ":labels"].astype(np.int64)
    names = z[tag + ":chrom_names"]
    start = z[tag + ":start"].astype(np.int64)
    chr_pos = pd.DataFrame({"chrom": names[z[tag + ":chrom"]], "start": start, "end": start + 1, "strand": "+"})
    return data_local, z[tag + ":prob"], chr_pos


@pytest.mark.parametrize("tag,n_class,kmers", TAGS)
def test_evaluator_matches_reference_outputs(tag, n_class, kmers):
    from mural_b200.evaluation import Evaluator, calc_avg_prob
    z = np.load(os.path.join(GOLD, "eval_kat.npz"))
    data_local, prob, chr_pos = _frames(z, tag, n_class)
    tol = 2e-5 if prob.dtype == np.float32 else 1e-7
    lines = []
    E = Evaluator(data_local, prob, n_class, printer=lambda *a: lines.append(a))
    E.evaluate_kmer(kmers)
    for k in kmers:
        assert np.allclose(E.metrics["kmer%d" % k], z["%s:kmer%d" % (tag, k)], rtol=0, atol=tol, equal_nan=True), (k, E.metrics["kmer%d" % k])
    assert [a[0] for a in lines if isinstance(a[0], str) and "mer correlation" in a[0]] == ["%dmer correlation - all: " % k for k in kmers]
    assert np.allclose(calc_avg_prob(E.ed, n_class), z[tag + ":avg_prob"], rtol=0, atol=1e-7)
    E.evaluate_regional_score(len(prob), kmers[:2])
    ref_score, ref_regions = z[tag + ":regional_score"]
    assert ("n_regions:", int(ref_regions)) in lines
    assert abs(E.metrics["score"] - ref_score) < 20 * tol * max(1.0, ref_score), (E.metrics["score"], ref_score)
    assert np.allclose(E.metrics["corr_list"], z[tag + ":regional_corr_list"], rtol=0, atol=10 * tol, equal_nan=True)
    # rows arrive in prediction order (unsorted chromosomes by NAME): the evaluator orders them like sort_values(['chrom','start'])
    perm = np.random.default_rng(0).permutation(len(prob))
    E2 = Evaluator(data_local.iloc[perm].reset_index(drop=True), prob[perm], n_class, printer=lambda *a: None)
    E2.evaluate_regional_corr(chr_pos.iloc[perm].reset_index(drop=True))
    for w in (100000, 500000):
        wtol = 2e-4 if prob.dtype == np.float32 else tol
        assert np.allclose(E2.metrics["window%d" % w], z["%s:window%d" % (tag, w)], rtol=0, atol=wtol, equal_nan=True), (w, E2.metrics["window%d" % w])


def test_group_tables_exact_and_reproducible():
    """Counts are exact, fixed-point sums are independent of the atomics' order: two runs give identical tables; the means
    agree with float64 numpy sums to 2^-36; every region of evaluate_regional_score equals a launch on its slice."""
    from mural_b200.evaluation import SCALE, EvalData, kmer_group_table
    rng = np.random.default_rng(5)
    n, K = 300000, 4
    flank = rng.integers(0, 5, (n, 15)).astype(np.int64)
    labels = rng.integers(0, K, n)
    prob = rng.dirichlet(np.ones(K), n)
    ed = EvalData(torch.from_numpy(flank).cuda(), torch.from_numpy((labels << 1).astype(np.int32)).cuda(), torch.from_numpy(prob).cuda())
    for k in (3, 5, 7):
        t1, t2 = kmer_group_table(ed, k), kmer_group_table(ed, k)
        assert np.array_equal(t1, t2)
        ids, obs, pred = EN.kmer_group_means(flank, labels, prob, k, K)
        tab = t1[0]
        assert np.array_equal(np.nonzero(tab[:, 0])[0], ids)
        cnt = tab[ids, 0].astype(np.float64)
        assert int(tab[:, 0].sum()) == n
        assert np.array_equal(tab[ids, 1:1 + K] / cnt[:, None], obs)
        assert np.abs(tab[ids, 1 + K:] / SCALE / cnt[:, None] - pred).max() < 2.0 ** -36
    reg = kmer_group_table(ed, 5, 10000)
    assert reg.shape[0] == 30
    sl = EvalData(ed.flank[70000:80000], ed.meta[70000:80000], ed.prob[70000:80000])
    assert np.array_equal(reg[7], kmer_group_table(sl, 5)[0])


def test_window_runs_vs_oracle_large_and_edges():
    from mural_b200.evaluation import EvalData, window_table
    rng = np.random.default_rng(9)
    K = 4
    for n in (1, 2, 2047, 2048, 2049, 500000):
        chrom = np.sort(rng.integers(0, 5, n))
        start = np.concatenate([np.sort(rng.integers(0, 40_000_000, int((chrom == c).sum()))) for c in range(5)]).astype(np.int64)
        labels = rng.integers(0, K, n)
        prob = rng.dirichlet(np.ones(K), n)
        meta = torch.from_numpy(((labels << 1) | (chrom << 8)).astype(np.int32)).cuda()
        ed = EvalData(torch.zeros((n, 1), dtype=torch.int64, device="cuda"), meta, torch.from_numpy(prob).cuda(),
                      start=torch.from_numpy(start.astype(np.int32)).cuda())
        for window in (1, 1000, 100000, 2_000_000_000):
            ref = EN.window_table(chrom, start, labels, prob, window, K)
            got = window_table(ed, window)
            assert got.shape == ref.shape, (n, window, got.shape, ref.shape)
            assert np.array_equal(got[:, :K], ref[:, :K])
            assert np.abs(got[:, K:] - ref[:, K:]).max() < 1e-10
    # explicit order: shuffled storage, sorted traversal
    perm = rng.permutation(n)
    inv = np.argsort(perm)
    ed_p = EvalData(ed.flank, meta[torch.from_numpy(perm).cuda()], ed.prob[torch.from_numpy(perm).cuda()],
                    start=ed.start[torch.from_numpy(perm).cuda()])
    got = window_table(ed_p, 100000, torch.from_numpy(inv).cuda())
    assert np.array_equal(got, window_table(ed, 100000))


def test_metric_errors():
    from mural_b200.evaluation import EvalData, kmer_group_table
    ed = EvalData(torch.full((10, 15), 7, dtype=torch.int64, device="cuda"), torch.zeros(10, dtype=torch.int32, device="cuda"),
                  torch.full((10, 4), .25, dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        kmer_group_table(ed, 3)           # code outside 0..4
    ed.flank.zero_()
    with pytest.raises(ValueError):
        kmer_group_table(ed, 17)          # k-mer longer than the local columns
    assert kmer_group_table(ed, 3)[0][0, 0] == 10
