#!/bin/bash
# round 2, call A: baseline of the round-1 binary on this box + ncu evidence for the shipped persistent K1
set -x
cd "$GRAFT_REPO_ROOT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_n256.json 2> gpurun_out/r02a_bench_n256.err
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --layout packed8 --n-bas 384 --n-spin 1 > gpurun_out/r02a_bench_n384p.json 2>> gpurun_out/r02a_bench_n256.err
timeout 300 $B > gpurun_out/r02a_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02a_launches_n256_uhf.csv $B > gpurun_out/r02a_ncu_l.log 2>&1
timeout 300 $B > gpurun_out/r02a_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_full_tma -s 4 -c 2 -f -o gpurun_out/r02a_k1_tma_n256_uhf $B > gpurun_out/r02a_ncu_full.log 2>&1
tail -3 gpurun_out/r02a_ncu_full.log
cat gpurun_out/r02a_bench_n256.json gpurun_out/r02a_bench_n384p.json | cut -c1-600
