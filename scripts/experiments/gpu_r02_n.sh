#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for ev in 0 1 0 1; do
SCFB_FULL_TMA_EVEN_WAVES=$ev python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 8 --rank 5 --steps 300 | tail -1 | sed "s/^/even=$ev /"
done
for ev in 0 1; do
SCFB_FULL_TMA_EVEN_WAVES=$ev python benchmarks/shard_bsplit.py --child --n-bas 256 --n-spin 2 --ranks 4 --rank 1 --steps 150 | tail -1 | sed "s/^/even=$ev /"
SCFB_FULL_TMA_EVEN_WAVES=$ev python benchmarks/shard_bsplit.py --child --n-bas 128 --n-spin 1 --ranks 1 --rank 0 --steps 400 | tail -1 | sed "s/^/even=$ev /"
SCFB_FULL_TMA_EVEN_WAVES=$ev python benchmarks/shard_bsplit.py --child --n-bas 128 --n-spin 2 --ranks 8 --rank 3 --steps 400 | tail -1 | sed "s/^/even=$ev /"
done
