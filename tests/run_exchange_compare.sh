# N=8 comparison of the two cross-GPU exchanges (run under: gpurun --gpus 8 -- bash tests/run_exchange_compare.sh)
R="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
p=29600
$R --master-port $p tests/mgpu_check.py p2p 2>&1 | grep -E "mgpu_check|Error|assert" | tail -3
for x in nccl p2p nccl p2p; do
  p=$((p+1))
  $R --master-port $p bench.py --gpus 8 --steps 200 --warmup 10 --exchange $x 2>/dev/null | grep "^{" > gpurun_out/x8_n256_$x.json
  python -c "import json; d=json.load(open('gpurun_out/x8_n256_$x.json')); print('n256 uhf', '$x', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value'],1))"
done
for x in nccl p2p; do
  p=$((p+1))
  $R --master-port $p bench.py --gpus 8 --layout packed8 --n-bas 384 --n-spin 1 --steps 100 --warmup 10 --exchange $x 2>/dev/null | grep "^{" > gpurun_out/x8_n384p_$x.json
  python -c "import json; d=json.load(open('gpurun_out/x8_n384p_$x.json')); print('n384 packed rhf', '$x', round(d['value'],1), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms'],4), round(d['e2e']['value'],1))"
done
