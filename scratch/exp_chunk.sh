for c in 32768 65536 131072; do MURAL_TC_CHUNK=$c timeout 200 python bench.py --steps 5 --warmup 2 --no-cpu-baseline > /tmp/b_$c.json 2>/tmp/b_$c.err; tail -2 /tmp/b_$c.err; python -c "
import json; d=json.load(open('/tmp/b_$c.json')); r=d['roofline']; print('chunk=$c', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), {k:round(v/5,2) for k,v in r['profile_ms'].items()})"; done
