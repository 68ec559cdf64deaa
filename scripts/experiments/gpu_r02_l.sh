#!/bin/bash
cd "$GRAFT_REPO_ROOT"
SCFB_FULL_TC=4 timeout 600 python -m pytest tests/test_jk_gpu.py tests/test_jk_fullsize_gpu.py -x -q -k "not packed" 2>&1 | tail -2
for tc in 8 4 8 4; do
SCFB_FULL_TC=$tc python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 8 --rank 5 --steps 200 | tail -1 | sed "s/^/tc=$tc /"
done
for tc in 8 4; do
SCFB_FULL_TC=$tc python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 4 --rank 1 --steps 100 | tail -1 | sed "s/^/tc=$tc /"
SCFB_FULL_TC=$tc python benchmarks/shard_bsplit.py --child --n-bas 128 --n-spin 1 --ranks 1 --rank 0 --steps 300 | tail -1 | sed "s/^/tc=$tc /"
SCFB_FULL_TC=$tc python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 1 --rank 0 --steps 40 | tail -1 | sed "s/^/tc=$tc /"
done
