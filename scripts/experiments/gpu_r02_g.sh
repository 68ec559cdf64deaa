#!/bin/bash
# multi-GPU parity: both exchanges (run with gpurun --gpus N)
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
R="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
p=29610
for x in p2p nccl; do
  p=$((p+1))
  $R --master-port $p tests/mgpu_check.py $x > gpurun_out/r02_mgpu_check_w${N}_$x.log 2>&1
  grep -E "mgpu_check|Error|assert|Traceback" gpurun_out/r02_mgpu_check_w${N}_$x.log | tail -5
done
for x in p2p nccl; do
  p=$((p+1))
  $R --master-port $p bench.py --gpus $N --steps 100 --warmup 10 --exchange $x 2>/dev/null | grep "^{" > gpurun_out/r02g_n256_w${N}_$x.json
  python -c "import json; d=json.load(open('gpurun_out/r02g_n256_w${N}_$x.json')); print('n256 uhf w$N', '$x', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value'],1))"
done
