#!/bin/bash
# round 2, single GPU: full GPU test suite, default bench line + reference arm, ncu evidence for the shipped binaries
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py > $O/r02_bench_default.json 2> $O/r02_bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/r02_bench_reference.json 2>> $O/r02_bench_default.err
timeout 300 python bench.py --n-bas 128 --n-spin 1 --no-extra --no-cpu-baseline --steps 200 > $O/r02_bench_n128_rhf.json 2>> $O/r02_bench_default.err
timeout 300 python bench.py --layout packed8 --n-bas 384 --n-spin 2 --no-extra --no-cpu-baseline --steps 50 > $O/r02_bench_n384_packed_uhf.json 2>> $O/r02_bench_default.err
timeout 300 python bench.py --layout packed8 --n-bas 384 --n-spin 1 --no-extra --no-cpu-baseline --steps 100 > $O/r02_bench_n384_packed_rhf.json 2>> $O/r02_bench_default.err
for n in 128 256; do timeout 300 python benchmarks/scf_loop.py --n-bas $n > $O/r02_scf_loop_n$n.json 2>&1; done
timeout 300 python benchmarks/scf_loop.py --n-bas 128 --unrestricted > $O/r02_scf_loop_n128_uhf.json 2>&1
timeout 300 python benchmarks/scf_loop.py --n-bas 256 --layout packed8 > $O/r02_scf_loop_n256_packed.json 2>&1
B="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extra"
timeout 300 $B > $O/r02_plain1.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_n256_uhf.csv $B > $O/r02_ncu1.log 2>&1
timeout 300 $B > $O/r02_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_full_tma -s 4 -c 2 -f -o $O/r02_k1_tma_n256_uhf $B > $O/r02_ncu2.log 2>&1
S="python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 8 --rank 5 --steps 3"
timeout 300 $S > $O/r02_plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:jk_full_tma -s 4 -c 2 -f -o $O/r02_k1_tma_n256_uhf_shard5of8 $S > $O/r02_ncu3.log 2>&1
for sp in 1 2; do
P="python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-extra --layout packed8 --n-bas 384 --n-spin $sp"
timeout 300 $P > $O/r02_plain4_$sp.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jk_packed_kernel -s 3 -c 1 -f -o $O/r02_k2_n384_spin$sp $P > $O/r02_ncu4_$sp.log 2>&1
done
timeout 300 $P > $O/r02_plain5.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r02_launches_n384_packed_uhf.csv $P > $O/r02_ncu5.log 2>&1
L="python benchmarks/scf_loop.py --n-bas 128"
timeout 300 $L > $O/r02_plain6.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 400 --csv --log-file $O/r02_launches_scf_loop_n128.csv $L > $O/r02_ncu6.log 2>&1
ls -la $O/r02_*.ncu-rep | awk '{print $5, $9}'
head -c 700 $O/r02_bench_default.json; echo; cat $O/r02_scf_loop_n128.json | cut -c1-400
