for d in 0 4 2 6 8 1 15; do MURAL_NO_LATTICE=1 MURAL_TC_DBG=$d timeout 120 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --sites-per-step 262144 > /tmp/b_$d.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/b_$d.json')); r=d['roofline']; print('dbg=$d', round(d['value']/1e6,2), {k:round(v/3,2) for k,v in r['profile_ms'].items() if 'stage' in k})"; done
