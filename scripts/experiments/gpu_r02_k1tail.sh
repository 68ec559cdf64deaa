#!/bin/bash
# K1 tail-split items: parity first, then the timeline with whole items / tail split / evict-first.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jk_gpu.py tests/test_jk_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
SCFB_FULL_BSPLIT=2 timeout 600 python -m pytest tests/test_jk_gpu.py -m gpu -x -q 2>&1 | tail -2
SCFB_FULL_BSPLIT=4 timeout 600 python -m pytest tests/test_jk_fullsize_gpu.py -m gpu -x -q -k "256" 2>&1 | tail -2
T=selfconsistentfield.jl_b200
{
echo "=== whole items (SCFB_FULL_TAIL=0), 1/8 shard"
SCFB_FULL_TAIL=0 timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 8 --rank 5 --builds 2
echo "=== tail split (default), 1/8 shard"
timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 8 --rank 5 --builds 2
echo "=== tail split + evict-first, 1/8 shard"
SCFB_LIB=$PWD/$T/libscf_b200_trace_ef.so timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 8 --rank 5 --builds 2
echo "=== tail split 20 slabs, 1/8 shard"
SCFB_FULL_TAIL=20 timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 8 --rank 5 --builds 2
echo "=== tail split, full tensor"
timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 1 --rank 0 --builds 2
echo "=== tail split + evict-first, full tensor"
SCFB_LIB=$PWD/$T/libscf_b200_trace_ef.so timeout 300 python benchmarks/k1_trace.py --n-bas 256 --ranks 1 --rank 0 --builds 2
} > gpurun_out/k1_tail.log 2>&1
grep -E "===|--- build" gpurun_out/k1_tail.log
