#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for sp in 1 2; do
B="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --layout packed8 --n-bas 384 --n-spin $sp"
timeout 300 $B > gpurun_out/r02e_plain_s$sp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_packed_kernel -s 3 -c 1 -f -o gpurun_out/r02e_k2_s$sp $B > gpurun_out/r02e_ncu_s$sp.log 2>&1
tail -1 gpurun_out/r02e_ncu_s$sp.log
done
