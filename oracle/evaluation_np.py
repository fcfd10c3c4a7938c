"""TEST INFRASTRUCTURE ONLY (oracle/): numpy restatement of the reference's per-epoch validation metrics.

Follows MuRaL/evaluation/evaluation.py: freq_kmer_comp_multi (:48-67), corr_calc_sub (:124-193), calc_avg_prob (:196-203),
Evaluator.evaluate_regional_score (:545-566).  Pinned by tests/golden/eval_kat.npz, which oracle/make_golden_eval.py writes
from the unmodified reference functions (corr_calc_sub needs `DataFrame.append`, removed in pandas 2: the generator restores
it as a thin `pd.concat` shim, which does not touch the arithmetic).

dtype rule the reference inherits from pandas for float32 columns (the uncalibrated softmax output): `groupby(...).mean()`
runs a Kahan-compensated sum IN float32, row by row (pandas/_libs/groupby.pyx group_mean), `Series.mean()` a float32 numpy
pairwise sum (core/nanops.py nanmean), and both return float32; correlations then run in float64 on those means.
`f32_means=True` reproduces exactly that (python loop: small inputs only); calibrated probabilities are float64 throughout.
The CUDA path accumulates exactly (fixed point) and is therefore compared with a tolerance of a few float32 ulps of the
means when the input is float32, and tightly when it is float64.
"""
import numpy as np


def pearson(x, y):
    """Series.corr / scipy.stats.pearsonr: NaN for fewer than 2 points or a constant column."""
    x = np.asarray(x, np.float64)
    y = np.asarray(y, np.float64)
    if len(x) < 2:
        return float("nan")
    xm, ym = x - x.mean(), y - y.mean()
    den = np.sqrt((xm * xm).sum() * (ym * ym).sum())
    return float((xm * ym).sum() / den) if den > 0 else float("nan")


def _group_mean(inv, n_groups, cnt, x, f32):
    """groupby(...).mean() of one column: float64 -> plain sums; float32 -> pandas' float32 Kahan loop in row order."""
    if not f32:
        return np.bincount(inv, weights=np.asarray(x, np.float64), minlength=n_groups) / cnt
    x = np.asarray(x, np.float32)
    sumx = np.zeros(n_groups, np.float32)
    comp = np.zeros(n_groups, np.float32)
    for g, v in zip(inv.tolist(), x):
        y = v - comp[g]
        t = sumx[g] + y
        comp[g] = (t - sumx[g]) - y
        sumx[g] = t
    return (sumx / cnt.astype(np.float32)).astype(np.float64)


def flank_columns(n_cols, k):
    """Column indices of us_d..us1, ds1..ds_d in the order-1 local matrix [us_R..us1, mid, ds1..ds_R] (:53-54)."""
    d, mid = k // 2, n_cols // 2
    return [mid - j for j in range(d, 0, -1)] + [mid + j for j in range(1, d + 1)]


def kmer_group_means(flank, labels, prob, k, n_class, f32_means=False):
    """Per observed k-mer context (sorted like pandas groups): mean of [label == i] and of prob_i (:60-64).
    Returns (group_ids, obs [G, n_class], pred [G, n_class])."""
    flank = np.asarray(flank)
    gid = np.zeros(len(flank), np.int64)
    for c in flank_columns(flank.shape[1], k):
        gid = gid * 5 + flank[:, c]
    ids, inv, cnt = np.unique(gid, return_inverse=True, return_counts=True)
    obs = np.zeros((len(ids), n_class))
    pred = np.zeros((len(ids), n_class))
    for i in range(n_class):
        obs[:, i] = np.bincount(inv, weights=(np.asarray(labels) == i).astype(np.float64), minlength=len(ids)) / cnt
        pred[:, i] = _group_mean(inv, len(ids), cnt, np.asarray(prob)[:, i], f32_means)
    return ids, obs, pred


def freq_kmer_comp_multi(flank, labels, prob, k, n_class, f32_means=False):
    _, obs, pred = kmer_group_means(flank, labels, prob, k, n_class, f32_means)
    return [pearson(obs[:, i], pred[:, i]) for i in range(n_class)]


def calc_avg_prob(labels, prob, n_class, f32_means=False):
    labels = np.asarray(labels)
    out = [float((labels == i).sum() / len(labels)) for i in range(n_class)]
    prob = np.asarray(prob)
    if f32_means:  # Series.mean of float32: numpy pairwise sum in float32, divided by the count
        out += [float(prob[:, i].astype(np.float32).sum(dtype=np.float32) / np.float32(len(labels))) for i in range(n_class)]
    else:
        out += [float(prob[:, i].astype(np.float64).mean()) for i in range(n_class)]
    return out


def window_table(chrom, start, labels, prob, window, n_class, f32_sums=False):
    """avg_obs / avg_pred of every run of consecutive rows sharing (chrom, start // window) (:139-171); rows as given
    (the caller has sorted them by chrom name and start, :526-528).  Returns [runs, 2 * n_class] (obs columns, then pred).
    The reference adds the frame's scalars one by one (:167): with float32 columns that is a plain sequential float32 sum
    divided in float32 (`f32_sums=True`), with float64 columns a float64 one."""
    chrom, start, labels = np.asarray(chrom), np.asarray(start), np.asarray(labels)
    key = np.asarray(start) // window
    new = np.ones(len(start), bool)
    new[1:] = (chrom[1:] != chrom[:-1]) | (key[1:] != key[:-1])
    run = np.cumsum(new) - 1
    cnt = np.bincount(run).astype(np.float64)
    tab = np.zeros((len(cnt), 2 * n_class))
    for i in range(n_class):
        tab[:, i] = np.bincount(run, weights=(labels == i).astype(np.float64)) / cnt
        if f32_sums:
            acc = np.zeros(len(cnt), np.float32)
            for r, v in zip(run.tolist(), np.asarray(prob)[:, i].astype(np.float32)):
                acc[r] += v
            tab[:, n_class + i] = (acc / cnt.astype(np.float32)).astype(np.float64)
        else:
            tab[:, n_class + i] = np.bincount(run, weights=np.asarray(prob)[:, i].astype(np.float64)) / cnt
    return tab


def corr_calc_sub(chrom, start, labels, prob, window, n_class, f32_sums=False):
    tab = window_table(chrom, start, labels, prob, window, n_class, f32_sums)
    if tab.shape[0] < 3:
        return [0] * n_class                                            # :186-188
    return [pearson(tab[:, i], tab[:, n_class + i]) for i in range(n_class)]


def regional_score(flank, labels, prob, valid_size, kmer_list, n_class, f32_means=False):
    """evaluate_regional_score (:545-566): (score, corr_list over region averages, n_regions)."""
    region_size = 10000 if valid_size > 10000 * 10 else valid_size // 10
    n_regions = valid_size // region_size
    score = 0.0
    avg = []
    for r in range(n_regions):
        sl = slice(region_size * r, region_size * (r + 1))
        for k in kmer_list[:2]:
            score += float(np.sum([(1 - c) ** 2 for c in freq_kmer_comp_multi(flank[sl], labels[sl], prob[sl], k, n_class, f32_means)]))
        avg.append(calc_avg_prob(labels[sl], prob[sl], n_class, f32_means))
    avg = np.asarray(avg)
    return score, [pearson(avg[:, i], avg[:, i + n_class]) for i in range(n_class)], n_regions
