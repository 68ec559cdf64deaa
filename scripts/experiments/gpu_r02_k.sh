#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_jk_packed_gpu.py tests/test_jk_fullsize_gpu.py -x -q -k packed 2>&1 | tail -2
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --layout packed8 --n-bas 384 --n-spin $1 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('n384 spin=$1', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"; }
run 1; run 2
for r in 0 3 7; do python benchmarks/shard_bsplit.py --child --layout packed8 --n-bas 384 --n-spin 1 --ranks 8 --rank $r --steps 50 | tail -1; done
