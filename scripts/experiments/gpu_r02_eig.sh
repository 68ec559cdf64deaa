#!/bin/bash
# Native Jacobi eigensolver (one CTA / 8-CTA cluster): parity, then SCF-loop timing against the cuSOLVER modes.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eig_jacobi_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_scf_gpu.py -m gpu -x -q -k "eigensolver or chain" 2>&1 | tail -5
{
for n in 128 64 168; do
  echo "== SCFB_EIG=native n=$n RHF"; SCFB_EIG=native timeout 300 python benchmarks/scf_loop.py --n-bas $n
done
echo "== SCFB_EIG=native n=128 UHF"; SCFB_EIG=native timeout 300 python benchmarks/scf_loop.py --n-bas 128 --unrestricted
echo "== SCFB_EIG=native n=256 RHF"; SCFB_EIG=native timeout 300 python benchmarks/scf_loop.py --n-bas 256
echo "== SCFB_EIG=syevd n=256 RHF"; SCFB_EIG=syevd timeout 300 python benchmarks/scf_loop.py --n-bas 256
echo "== SCFB_EIG=native n=64 RHF cluster"; SCFB_EIG_CLUSTER_MIN=32 SCFB_EIG=native timeout 300 python benchmarks/scf_loop.py --n-bas 64
echo "== sweeps n=128 RHF"; SCFB_VERBOSE=1 SCFB_EIG=native timeout 300 python benchmarks/scf_loop.py --n-bas 128 2>&1 | grep -E "sweeps" | tail -9 | tr '\n' ' '
echo
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/eig_launches.csv -k regex:jacobi_eig python benchmarks/scf_loop.py --n-bas 128 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/eig_launches.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; st=i+1; break
vi=hdr.index('Metric Value')
v=[float(r[vi].replace(',','')) for r in rows[st:] if len(r)>vi]
print("jacobi kernel us per launch (ncu):", [round(x/1e3) for x in v])
PY
} > gpurun_out/eig_loop.log 2>&1
grep -v Removing gpurun_out/eig_loop.log | cut -c1-470
