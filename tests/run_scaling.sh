run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%400)) "$@" 2>&1 | grep -E '^\{|mgpu_check|Error|error' | tail -2; }
run 8 tests/mgpu_check.py
for n in 4 8; do run $n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/scale_n256_g$n.json; cat gpurun_out/scale_n256_g$n.json | cut -c1-120; done
run 8 bench.py --gpus 8 --n-bas 512 --n-spin 1 --steps 30 --warmup 3 > gpurun_out/scale_n512_rhf_g8.json; cut -c1-150 gpurun_out/scale_n512_rhf_g8.json
for n in 2 4 8; do run $n bench.py --gpus $n --layout packed8 --n-bas 384 --n-spin 1 --steps 50 --warmup 5 > gpurun_out/scale_n384p_g$n.json; cut -c1-120 gpurun_out/scale_n384p_g$n.json; done
