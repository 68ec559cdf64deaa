#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for eig in sygvd syevd warm; do
for args in "--n-bas 128" "--n-bas 256" "--n-bas 128 --unrestricted"; do
SCFB_EIG=$eig timeout 300 python benchmarks/scf_loop.py $args 2>&1 | tail -1 | cut -c1-250 | sed "s/^/eig=$eig $args: /"
done
done
