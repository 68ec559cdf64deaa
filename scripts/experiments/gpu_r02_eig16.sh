#!/bin/bash
# 16-CTA cluster Jacobi (320 < n <= 512): parity, then packed8 SCF loops against Dsyevd
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eig_jacobi_gpu.py -m gpu -x -q 2>&1 | tail -6
{
echo "== default n=512 packed"; timeout 400 python benchmarks/scf_loop.py --n-bas 512 --layout packed8
echo "== native n=512 packed"; SCFB_EIG=native timeout 400 python benchmarks/scf_loop.py --n-bas 512 --layout packed8
} > gpurun_out/eig16b_loop.log 2>&1
grep -v Removing gpurun_out/eig16b_loop.log | cut -c1-480
