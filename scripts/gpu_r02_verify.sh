#!/bin/bash
# last check of the tree as committed: GPU tests, smoke(), default bench line, reference arm
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02_verify_bench.json 2> gpurun_out/r02_verify_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_verify_bench.json') if l.startswith('{')][0])
print('bench', round(d['value'], 1), round(d['ms_per_step'], 4), 'share', round(d['roofline']['kernel_share_of_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['value'], 1), 'launches', d['gpu_launches'])
print(d['verified'][:120])
for x in d.get('extra_configs', []): print('extra', x['config']['workload'][:45], round(x['value'], 1), 'frac', round(x['roofline']['frac'], 3), 'traffic', x['roofline']['traffic'])
print('cpu', round(d['cpu_baseline']['value'], 4), d['cpu_baseline']['cores'])
PY
grep real gpurun_out/r02_verify_bench.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) 2>&1 | cut -c1-200 | tail -5
