#!/bin/bash
# K1 timeline on a 1/8 shard, a 1/4 shard and the full tensor (tuning build with -DSCFB_K1_TRACE).
mkdir -p gpurun_out
{
timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 8 --rank 5 --builds 3
timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 4 --rank 1 --builds 2
timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 1 --rank 0 --builds 2
} > gpurun_out/k1_trace.log 2>&1
tail -5 gpurun_out/k1_trace.log
