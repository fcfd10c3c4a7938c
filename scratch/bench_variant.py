"""bench.py's dense predict leg against a library variant: MURAL_LIB=scratch/libmural_X.so python scratch/bench_variant.py"""
import os, sys, json, io, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mural_b200._lib as L
if os.environ.get("MURAL_LIB"): L.LIB_PATH = os.path.join(ROOT, os.environ["MURAL_LIB"])
sys.argv = ["bench.py", "--mode", os.environ.get("MODE", "auto"), "--steps", os.environ.get("STEPS", "6"), "--warmup", "3", "--no-sparse", "--no-pipeline", "--no-cpu-baseline", "--no-train", "--no-indel", "--no-sweep", "--no-eval"]
import bench
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads(buf.getvalue().strip().splitlines()[-1])
p = d["roofline"]["profile_ms"]; c = d["roofline"]["profile_count"]
bo = (d["config"].get("bf16_only") or {}).get("ms_per_step", float("nan"))
print(os.environ.get("MURAL_LIB", "product"), "%s %.3f ms, bf16-only %.3f ms | " % (d["config"]["mode"], d["ms_per_step"], bo) +
      ", ".join("%s %.3f" % (k, v / max(1, c[k])) for k, v in sorted(p.items(), key=lambda kv: -kv[1])[:6]) + " (ms per launch)")
