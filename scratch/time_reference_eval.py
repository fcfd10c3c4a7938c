"""Container-only: wall time of the unmodified reference's evaluation functions on the eval_kat fixture (24 000 sites)."""
import contextlib, importlib, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, pandas as pd
from oracle import ref_import
ref_import.install_stubs(); sys.path.insert(0, ref_import.REF_ROOT)
if not hasattr(sys.modules["jax"], "Array"): sys.modules["jax"].Array = type("Array", (), {})
if not hasattr(pd.DataFrame, "append"):
    pd.DataFrame.append = lambda self, other, **kw: other.copy() if len(self) == 0 else pd.concat([self, other], **kw)
_ror = pd.Series.__ror__
pd.Series.__ror__ = lambda self, other: _ror(self, np.asarray(other) if isinstance(other, list) else other)
ev = importlib.import_module("MuRaL.evaluation.evaluation")
z = np.load(os.path.join(ROOT, "tests", "golden", "eval_kat.npz"))
tag, K = "snv_f32", 4
flank = z[tag + ":flank"].astype(np.int64); n = len(flank)
cols = ["us%d" % i for i in range(7, 0, -1)] + ["mid"] + ["ds%d" % i for i in range(1, 8)]
dl = pd.DataFrame(flank, columns=cols); dl["mut_type"] = z[tag + ":labels"].astype(np.int64)
E = ev.Evaluator(dl, z[tag + ":prob"], K, printer=lambda *a: None)
t0 = time.perf_counter(); E.evaluate_kmer([3, 5, 7]); t1 = time.perf_counter(); E.evaluate_regional_score(n, [3, 5]); t2 = time.perf_counter()
names = z[tag + ":chrom_names"]; start = z[tag + ":start"].astype(np.int64)
chr_pos = pd.DataFrame({"chrom": names[z[tag + ":chrom"]], "start": start, "end": start + 1, "strand": "+"})
with contextlib.redirect_stdout(io.StringIO()):
    E.evaluate_regional_corr(chr_pos)
t3 = time.perf_counter()
print("reference Evaluator on %d sites: evaluate_kmer %.3f s, evaluate_regional_score %.3f s, evaluate_regional_corr %.2f s -> %.0f sites/s overall" %
      (n, t1 - t0, t2 - t1, t3 - t2, n / (t3 - t0)))
