#!/bin/bash
# round-2 closing run on one B200: GPU suite, smoke, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gpu_tests.txt 2>&1; tail -2 gpurun_out/r02g_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.txt 2>&1; tail -3 gpurun_out/r02g_smoke.txt
python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02g_bench.json").read().strip().splitlines()[-1])
print(d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["roofline"]["frac"], d["config"].get("bf16_only", {}).get("ms_per_step"))
PY
