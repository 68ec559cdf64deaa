#!/bin/bash
# round 2, last pass on the tree as committed: GPU tests, smoke(), default bench line, and the captures of the
# eigensolver kernel in its final form (launch list of an SCF loop + ncu --set full of two K5b launches)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 3 --no-extra > $O/r02_verify_bench.json 2> $O/r02_verify_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/r02_verify_bench.json') if l.startswith('{')][0])
print('bench', round(d['value'], 1), round(d['ms_per_step'], 4), 'share', round(d['roofline']['kernel_share_of_step'], 4), 'frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value'], 1), 'cpu', round(d['cpu_baseline']['value'], 4))
PY
for n in 128 256; do timeout 300 python benchmarks/scf_loop.py --n-bas $n > $O/r02_scf_loop_n$n.json 2>&1; done
timeout 300 python benchmarks/scf_loop.py --n-bas 128 --unrestricted > $O/r02_scf_loop_n128_uhf.json 2>&1
L="python benchmarks/scf_loop.py --n-bas 128"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $O/r02_launches_scf_loop_n128.csv $L > $O/r02_ncu6.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:jacobi_eig_cluster -s 10 -c 2 -f -o $O/r02_k5b_jacobi_n128 $L > $O/r02_ncu7.log 2>&1
ls -la $O/r02_k5b_jacobi_n128.ncu-rep | awk '{print $5, $9}'
for f in $O/r02_scf_loop_n128.json $O/r02_scf_loop_n128_uhf.json $O/r02_scf_loop_n256.json; do grep -h '^{' $f | cut -c1-330; done
