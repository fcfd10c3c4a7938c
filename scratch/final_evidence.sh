#!/bin/bash
# round-2 final evidence on one B200: ncu launch list + full capture of the stage kernels, GPU suite, smoke, bench line
mkdir -p gpurun_out
B="--no-sparse --no-pipeline --no-cpu-baseline --no-train --no-indel --no-sweep --no-eval"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/r02f_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stage_tc -c 12 -f -o gpurun_out/r02f_stage python bench.py --steps 1 --warmup 1 --sites-per-step 524288 --mode bf16 $B > gpurun_out/r02f_stage_bench.log 2>&1
ncu --set full --clock-control none -k "regex:k_tail|k_local_mlp_tc|k_dense_tables|k_edge_pool|k_lattice_pool|k_stem_gather" -c 8 -f -o gpurun_out/r02f_aux python bench.py --steps 1 --warmup 1 --sites-per-step 524288 --mode bf16 $B > gpurun_out/r02f_aux_bench.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_gpu_tests.txt 2>&1; tail -3 gpurun_out/r02f_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.txt 2>&1; tail -3 gpurun_out/r02f_smoke.txt
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02f_bench.json").read().strip().splitlines()[-1])
print(d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["roofline"]["frac"], d["config"].get("bf16_only", {}).get("ms_per_step"))
PY
