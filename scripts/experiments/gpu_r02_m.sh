#!/bin/bash
cd "$GRAFT_REPO_ROOT"
SCFB_PK_UHF_KC8=1 timeout 300 python -m pytest tests/test_jk_packed_gpu.py tests/test_jk_fullsize_gpu.py -x -q -k packed 2>&1 | tail -2
run() { SCFB_PK_UHF_KC8=$1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --layout packed8 --n-bas 384 --n-spin $2 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kc8=$1 spin=$2', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"; }
run 0 2; run 1 2; run 0 2; run 1 2; run 0 1
