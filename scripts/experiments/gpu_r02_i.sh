#!/bin/bash
# multi-GPU: single-process multi handle tests + checks; torchrun parity both exchanges; scaling lines
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
timeout 600 python -m pytest tests/test_multi_handle_gpu.py -x -q 2>&1 | tail -4
timeout 900 python tests/multi_handle_check.py $N > gpurun_out/r02_multi_handle_w$N.log 2>&1; tail -5 gpurun_out/r02_multi_handle_w$N.log | cut -c1-400
bash scripts/gpu_r02_g.sh $N
