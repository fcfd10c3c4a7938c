#!/bin/bash
# usage: bench_short.sh [bench args]  -> one line: sites/s, ms/step, per-kernel ms per step
python bench.py --no-train --no-indel --no-sweep --no-eval --no-pipeline --no-cpu-baseline --no-sparse "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%.2f M sites/s  %.3f ms/step  frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['frac']), d['config'].get('bf16_only') and round(d['config']['bf16_only']['ms_per_step'],3))
print({k: round(v/d['steps'],3) for k,v in d['roofline']['profile_ms'].items() if v>0.5})"
