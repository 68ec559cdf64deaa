#!/bin/bash
# 16-CTA cluster Jacobi (320 < n <= 448): parity, then the n_bf = 384 packed8 SCF loop against Dsyevd
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eig_jacobi_gpu.py -m gpu -x -q 2>&1 | tail -6
{
echo "== default (syevd) n=384 packed"; timeout 300 python benchmarks/scf_loop.py --n-bas 384 --layout packed8
echo "== native n=384 packed"; SCFB_EIG=native timeout 300 python benchmarks/scf_loop.py --n-bas 384 --layout packed8
} > gpurun_out/eig16_loop.log 2>&1
grep -v Removing gpurun_out/eig16_loop.log | cut -c1-480
