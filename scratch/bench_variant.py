"""bench.py's dense predict leg against a library variant: MURAL_LIB=scratch/libmural_X.so python scratch/bench_variant.py"""
import os, sys, json, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mural_b200._lib as L
if os.environ.get("MURAL_LIB"): L.LIB_PATH = os.path.join(ROOT, os.environ["MURAL_LIB"])
sys.argv = ["bench.py", "--steps", "6", "--warmup", "3", "--no-sparse", "--no-pipeline", "--no-cpu-baseline", "--no-train", "--no-indel", "--no-sweep", "--no-eval"]
import bench
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads(buf.getvalue().strip().splitlines()[-1])
p = d["roofline"]["profile_ms"]; c = d["roofline"]["profile_count"]
print(os.environ.get("MURAL_LIB", "product"), "auto %.3f ms, bf16 %.3f ms | tail %.3f dense_tables %.3f (ms per launch)" % (
    d["ms_per_step"], d["config"]["bf16_only"]["ms_per_step"], p["(tail::k_tail)"] / c["(tail::k_tail)"], p["k_dense_tables<32>"] / c["k_dense_tables<32>"]))
