#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for nb in 256 128; do
timeout 300 python bench.py --n-bas $nb --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02f_err_$nb.log | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n=$nb', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], d['clocks']['sm_mhz'])"
done
tail -3 gpurun_out/r02f_err_256.log
