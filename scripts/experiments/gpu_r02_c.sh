#!/bin/bash
cd "$GRAFT_REPO_ROOT"
B="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --layout packed8 --n-bas 384 --n-spin 1"
timeout 300 $B > gpurun_out/r02c_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_packed_kernel -s 3 -c 1 -f -o gpurun_out/r02c_k2_n384_rhf $B > gpurun_out/r02c_ncu.log 2>&1
tail -2 gpurun_out/r02c_ncu.log
