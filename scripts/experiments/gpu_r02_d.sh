#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in 0 1 2 3; do
  echo "== variant $v tests"
  SCFB_PACKED_VARIANT=$v timeout 600 python -m pytest tests/test_jk_packed_gpu.py -x -q 2>&1 | tail -2
done
SCFB_PACKED_VARIANT=0 timeout 600 python -m pytest tests/test_jk_fullsize_gpu.py -x -q -k packed 2>&1 | tail -2
run() { # lib variant spin
  L=""; [ "$1" = r1 ] && L="$PWD/selfconsistentfield.jl_b200/libscf_b200_r1.so"
  SCFB_LIB=$L SCFB_PACKED_VARIANT=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --layout packed8 --n-bas 384 --n-spin $3 2>/dev/null | grep "^{" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lib=$1 variant=$2 spin=$3', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'frac', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])"
}
run r1 0 1
for v in 0 1 2 3; do run new $v 1; done
run r1 0 2
for v in 0 1; do run new $v 2; done
