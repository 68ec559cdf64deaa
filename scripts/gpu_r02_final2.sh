#!/bin/bash
# round 2, single GPU, after the K1 tail split + Jacobi eigensolver: GPU test suite, bench lines, ncu evidence
# for the kernels that changed since gpu_r02_final1.sh (K1, the SCF-loop launch list, the eigensolver).
# K2 (jk_packed_kernel) is byte-identical to the binary of gpu_r02_final1.sh: its captures stay.
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err
timeout 300 python bench.py --n-bas 128 --n-spin 1 --no-extra --no-cpu-baseline --steps 200 > $O/r02_bench_n128_rhf.json 2>> $O/r02_bench_default.err
for n in 64 128 168 256; do timeout 300 python benchmarks/scf_loop.py --n-bas $n > $O/r02_scf_loop_n$n.json 2>&1; done
timeout 300 python benchmarks/scf_loop.py --n-bas 128 --unrestricted > $O/r02_scf_loop_n128_uhf.json 2>&1
timeout 300 python benchmarks/scf_loop.py --n-bas 256 --layout packed8 > $O/r02_scf_loop_n256_packed.json 2>&1
B="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extra"
timeout 300 $B > $O/r02_plain1.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_n256_uhf.csv $B > $O/r02_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_full_tma -s 4 -c 2 -f -o $O/r02_k1_tma_n256_uhf $B > $O/r02_ncu2.log 2>&1
S="python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 8 --rank 5 --steps 3"
timeout 300 $S > $O/r02_plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:jk_full_tma -s 4 -c 2 -f -o $O/r02_k1_tma_n256_uhf_shard5of8 $S > $O/r02_ncu3.log 2>&1
L="python benchmarks/scf_loop.py --n-bas 128"
timeout 300 $L > $O/r02_plain6.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r02_launches_scf_loop_n128.csv $L > $O/r02_ncu6.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jacobi_eig_cluster -s 10 -c 2 -f -o $O/r02_k5b_jacobi_n128 $L > $O/r02_ncu7.log 2>&1
ls -la $O/r02_*.ncu-rep | awk '{print $5, $9}'
head -c 900 $O/r02_bench_default.json; echo; for f in $O/r02_scf_loop_n*.json; do grep -h '^{' $f | cut -c1-330; done
