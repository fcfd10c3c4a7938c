"""TEST INFRASTRUCTURE ONLY (oracle/): golden vectors of the validation metrics, from the unmodified reference.

Runs only in the build container (needs /root/reference).  Imports MuRaL/evaluation/evaluation.py, feeds it a synthetic
validation set shaped like `data_local` (order-1 local columns us7..us1, mid, ds1..ds7 + mut_type) and like the frame of
`evaluate_regional_corr` (chrom, start, end, strand, mut_type, prob*), and stores inputs and the reference's outputs in
tests/golden/eval_kat.npz; checks oracle/evaluation_np.py against them on the way.

`corr_calc_sub` relies on two pandas 1.x behaviours that pandas >= 2 / 3 removed (the reference pins pandas 1.x,
environment.yml): DataFrame.append, restored here as a pd.concat shim, and `list | Series` in its ">50 % zeros" warning
(:178), restored by converting the list to an array first.  No arithmetic is affected.

    python -m oracle.make_golden_eval
"""
import importlib
import io
import contextlib
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import evaluation_np as EN  # noqa: E402
from oracle import ref_import  # noqa: E402


def synth(n, n_class, seed, f32):
    rng = np.random.default_rng(seed)
    R = 7
    flank = rng.integers(0, 4, (n, 2 * R + 1)).astype(np.int64)
    flank[rng.random((n, 2 * R + 1)) < 0.002] = 4                       # a few N
    flank[:, R] = 0
    # rates that depend on the 5-mer context so the correlations are informative
    w = rng.normal(0, .6, (5, 5, 5, 5, n_class - 1))
    logit = np.concatenate([np.full((n, 1), 3.0), w[flank[:, R - 2], flank[:, R - 1], flank[:, R + 1], flank[:, R + 2]]], 1)
    p_true = np.exp(logit) / np.exp(logit).sum(1, keepdims=True)
    labels = np.array([rng.choice(n_class, p=p) for p in p_true]).astype(np.int64)
    noise = np.exp(rng.normal(0, .3, p_true.shape))
    prob = p_true * noise
    prob /= prob.sum(1, keepdims=True)
    prob = prob.astype(np.float32) if f32 else prob.astype(np.float64)
    chrom_names = np.array(["chr10", "chr2", "chrX"])                  # string sort order differs from numeric order
    chrom = np.sort(rng.integers(0, 3, n))
    start = np.concatenate([np.sort(rng.integers(0, 3_000_000, (chrom == c).sum())) for c in range(3)])
    return flank, labels, prob, chrom_names, chrom, start


def main():
    assert ref_import.available(), "needs /root/reference"
    ref_import.install_stubs()
    sys.path.insert(0, ref_import.REF_ROOT)
    if not hasattr(sys.modules["jax"], "Array"):       # scipy's array-API dispatch probes jax.Array on the import stub
        sys.modules["jax"].Array = type("Array", (), {})
    if not hasattr(pd.DataFrame, "append"):
        pd.DataFrame.append = lambda self, other, **kw: other.copy() if len(self) == 0 else pd.concat([self, other], **kw)
    _ror = pd.Series.__ror__
    pd.Series.__ror__ = lambda self, other: _ror(self, np.asarray(other) if isinstance(other, list) else other)
    ev = importlib.import_module("MuRaL.evaluation.evaluation")
    out = {}
    for tag, n, n_class, f32 in (("snv_f32", 24000, 4, True), ("snv_f64", 24000, 4, False), ("indel_f32", 12000, 8, True)):
        flank, labels, prob, chrom_names, chrom, start = synth(n, n_class, 11 + n_class + int(f32), f32)
        R = flank.shape[1] // 2
        cols = ["us%d" % i for i in range(R, 0, -1)] + ["mid"] + ["ds%d" % i for i in range(1, R + 1)]
        data_local = pd.DataFrame(flank, columns=cols)
        data_local["mut_type"] = labels
        lines = []
        E = ev.Evaluator(data_local, prob, n_class, printer=lambda *a: lines.append(a))
        kmers = [2, 4, 6] if n_class == 8 else [3, 5, 7]
        for k in kmers:
            ref = ev.freq_kmer_comp_multi(E.data_and_prob, k, n_class)
            got = EN.freq_kmer_comp_multi(flank, labels, prob, k, n_class, f32_means=f32)
            assert np.allclose(ref, got, rtol=0, atol=1e-12, equal_nan=True), (tag, k, ref, got)
            out["%s:kmer%d" % (tag, k)] = np.asarray(ref, np.float64)
        ref = ev.calc_avg_prob(E.data_and_prob, n_class)
        got = EN.calc_avg_prob(labels, prob, n_class, f32_means=f32)
        assert np.allclose(np.asarray(ref, np.float64), got, rtol=0, atol=1e-15), (tag, ref, got)
        out[tag + ":avg_prob"] = np.asarray(ref, np.float64)
        # evaluate_regional_score prints n_regions, corr_list, score
        E.evaluate_regional_score(n, kmers[:2])
        score, corr_list, n_regions = EN.regional_score(flank, labels, prob, n, kmers, n_class, f32_means=f32)
        assert abs(E.metrics["score"] - score) < 1e-9 * max(1, abs(score)), (tag, E.metrics["score"], score)
        ref_corr = [a for a in lines if isinstance(a[0], str) and a[0].startswith("corr_list")][0][1]
        assert np.allclose(ref_corr, corr_list, rtol=0, atol=1e-10, equal_nan=True), (tag, ref_corr, corr_list)
        out[tag + ":regional_score"] = np.array([E.metrics["score"], n_regions], np.float64)
        out[tag + ":regional_corr_list"] = np.asarray(ref_corr, np.float64)
        # corr_calc_sub on the frame evaluate_regional_corr builds (:522-528): sorted by chrom NAME, then start
        df = pd.DataFrame({"chrom": chrom_names[chrom], "start": start, "end": start + 1, "strand": "+", "mut_type": labels})
        for i in range(n_class):
            df["prob%d" % i] = prob[:, i]
        df.sort_values(["chrom", "start"], inplace=True)
        df.reset_index(drop=True, inplace=True)
        for window in (100000, 500000):
            with contextlib.redirect_stdout(io.StringIO()):
                ref = ev.corr_calc_sub(df, window, ["prob%d" % i for i in range(n_class)])
            order = np.lexsort((start, chrom_names[chrom]))
            got = EN.corr_calc_sub(chrom[order], start[order], labels[order], prob[order], window, n_class, f32_sums=f32)
            assert np.allclose(np.asarray(ref, np.float64), got, rtol=0, atol=1e-9, equal_nan=True), (tag, window, ref, got)
            out["%s:window%d" % (tag, window)] = np.asarray(ref, np.float64)
        for nm, v in (("flank", flank.astype(np.int8)), ("labels", labels.astype(np.int8)), ("prob", prob), ("chrom", chrom.astype(np.int8)),
                      ("start", start.astype(np.int32)), ("chrom_names", chrom_names)):
            out[tag + ":" + nm] = v
        print(tag, "reference == oracle:", {k.split(":")[1]: np.round(v, 4).tolist() for k, v in out.items()
                                            if k.startswith(tag + ":kmer") or k.startswith(tag + ":window")})
    path = os.path.join(ROOT, "tests", "golden", "eval_kat.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
