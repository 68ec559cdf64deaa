#!/bin/bash
# round 2, N GPUs, final binary: parity at world N over the fused peer-memory exchange + the scaling line
cd "$GRAFT_REPO_ROOT"
N=${1:-8}
O=gpurun_out
R="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$R --master-port 29811 tests/mgpu_check.py p2p > $O/r02_mgpu_check_w${N}_p2p.log 2>&1
grep -E "mgpu_check|Error|assert|Traceback" $O/r02_mgpu_check_w${N}_p2p.log | tail -3
$R --master-port 29812 bench.py --gpus $N --steps 100 --warmup 10 2>$O/r02_scale_g$N.err | grep "^{" > $O/r02_scale_n256_g$N.json
python - <<PY
import json
d = json.load(open("$O/r02_scale_n256_g$N.json"))
print(round(d['value'],1), round(d['ms_per_step'],4), 'k', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['config']['exchange'], d['verified'][:100])
for x in d.get('extra_configs', []): print('    ', x['config']['workload'][:42], round(x['value'],1), round(x['ms_per_step'],4), 'k', round(x['roofline']['kernel_ms'],4), 'frac', round(x['roofline']['frac'],3), 'e2e', round(x['e2e']['value'],1))
PY
tail -2 $O/r02_scale_g$N.err
