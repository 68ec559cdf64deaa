#!/bin/bash
cd "$GRAFT_REPO_ROOT"
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r02h_bench_default.json 2> gpurun_out/r02h_bench_default.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02h_bench_default.json') if l.startswith('{')][0])
print('head', round(d['value'],1), round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'share', round(d['roofline']['kernel_share_of_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
print(d['verified'])
for x in d.get('extra_configs', []): print('extra', x['config']['workload'], round(x['value'],1), round(x['ms_per_step'],4), 'frac', round(x['roofline']['frac'],3), 'e2e', round(x['e2e']['value'],1))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['cpu_baseline']['sample'][:120], d['cpu_baseline']['onepass_openmp_c_value'])
PY
tail -4 gpurun_out/r02h_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) 2>&1 | cut -c1-400 | tail -5
free -g | head -2; nproc
