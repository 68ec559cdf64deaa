#!/bin/bash
# round 2, N GPUs of one box: parity at world N (both exchanges), the single-process multi-GPU handle, scaling lines
cd "$GRAFT_REPO_ROOT"
N=${1:-2}
O=gpurun_out
R="timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
p=$((29700 + N * 10))
timeout 600 python -m pytest tests/test_multi_handle_gpu.py -x -q 2>&1 | tail -2
for x in p2p nccl; do
  p=$((p+1))
  $R --master-port $p tests/mgpu_check.py $x > $O/r02_mgpu_check_w${N}_$x.log 2>&1
  grep -E "mgpu_check|Error|assert|Traceback" $O/r02_mgpu_check_w${N}_$x.log | tail -3
done
timeout 900 python tests/multi_handle_check.py $N > $O/r02_multi_handle_w$N.log 2>&1; tail -4 $O/r02_multi_handle_w$N.log | cut -c1-330
p=$((p+1))
$R --master-port $p bench.py --gpus $N --steps 100 --warmup 10 2>$O/r02_scale_g$N.err | grep "^{" > $O/r02_scale_n256_g$N.json
p=$((p+1))
$R --master-port $p bench.py --gpus $N --steps 100 --warmup 10 --exchange nccl 2>>$O/r02_scale_g$N.err | grep "^{" > $O/r02_scale_n256_g${N}_nccl.json
python - <<PY
import json
for f in ["$O/r02_scale_n256_g$N.json", "$O/r02_scale_n256_g${N}_nccl.json"]:
    try: d = json.load(open(f))
    except Exception as e: print(f, "unreadable", e); continue
    print(f, round(d['value'],1), round(d['ms_per_step'],4), 'k', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value'],1), d['config']['exchange'])
    for x in d.get('extra_configs', []): print('    ', x['config']['workload'][:42], round(x['value'],1), round(x['ms_per_step'],4), 'k', round(x['roofline']['kernel_ms'],4), 'frac', round(x['roofline']['frac'],3), 'e2e', round(x['e2e']['value'],1))
PY
if [ "$N" = 8 ]; then
  timeout 600 python benchmarks/scf_loop.py --devices 8 --n-bas 384 --layout packed8 > $O/r02_scf_loop_n384_packed_dev8.json 2>&1; cut -c1-500 $O/r02_scf_loop_n384_packed_dev8.json | tail -1
  timeout 900 python benchmarks/scf_loop.py --devices 8 --n-bas 512 > $O/r02_scf_loop_n512_dev8.json 2>&1; cut -c1-500 $O/r02_scf_loop_n512_dev8.json | tail -1
fi
if [ "$N" = 2 ]; then
  for a in "--n-bas 128" "--n-bas 128 --unrestricted" "--n-bas 256"; do timeout 300 python benchmarks/scf_loop.py $a 2>&1 | tail -1 | cut -c1-420; done
fi
tail -3 $O/r02_scale_g$N.err
