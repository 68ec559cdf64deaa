#!/bin/bash
# K2 A/B: round-1 binary vs new kernel on the same box
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_jk_packed_gpu.py tests/test_jk_fullsize_gpu.py -x -q -k "packed" 2>&1 | tail -5
for lib in r1 new; do
  L=""; [ "$lib" = r1 ] && L="$PWD/selfconsistentfield.jl_b200/libscf_b200_r1.so"
  for sp in 1 2; do
    SCFB_LIB=$L timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --layout packed8 --n-bas 384 --n-spin $sp 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib=$lib n=384 spin=$sp', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
  done
done
