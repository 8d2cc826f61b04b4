#!/usr/bin/env bash
set -u
mkdir -p gpurun_out/r02_final4
timeout -k 10 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final4/bench_n1.json 2> gpurun_out/r02_final4/bench_n1.err; echo "bench rc=$?"
tail -3 gpurun_out/r02_final4/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_final4/bench_n1.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['extras'].items():
    if k.startswith('bm25'): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,dict)}, v.get('roofline',{}).get('frac'))
PY
