for c in 8192 16384 32768 65536; do MURAL_TC_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > /tmp/b_$c.json 2>/dev/null; python -c "
import json; d=json.load(open('/tmp/b_$c.json')); r=d['roofline']; print('chunk=$c', round(d['value']/1e6,2), {k:(round(v,1), r['profile_count'][k], round(1e3*v/r['profile_count'][k],1)) for k,v in r['profile_ms'].items()})"; done
