#!/bin/bash
# A/B of the tail-split plan with the shipped (non-trace) library: whole items vs default, several shard sizes.
mkdir -p gpurun_out
{
for rep in 1 2; do
for cfg in "256 8 5" "256 4 1" "256 2 1" "256 1 0" "128 1 0"; do
  set -- $cfg
  for tail in 0 1; do
    echo -n "n=$1 ranks=$2 tail=$tail: "
    SCFB_FULL_TAIL=$tail timeout 300 python benchmarks/shard_bsplit.py --n-bas $1 --ranks $2 --rank $3 --sweep auto --steps 60 | tail -1
  done
done
done
for tail in 0 1; do
  echo -n "n=512 ranks=8 rhf tail=$tail: "
  SCFB_FULL_TAIL=$tail timeout 600 python benchmarks/shard_bsplit.py --n-bas 512 --n-spin 1 --ranks 8 --rank 3 --sweep auto --steps 20 | tail -1
done
} > gpurun_out/k1_tail2.log 2>&1
cat gpurun_out/k1_tail2.log
