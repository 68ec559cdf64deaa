"""Multi-GPU parity check, one process per GPU (run under torchrun on the GPU box):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/mgpu_check.py
Every rank must return the FULL J/K (its shard's partial allreduced inside libscf_b200 with
NCCL), identical on all ranks and equal to the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import scf_b200  # noqa: E402
from oracle import scf_oracle as o  # noqa: E402
from oracle import synth  # noqa: E402


def main():
    rank, local_rank, world = scf_b200.distributed.env_ranks()
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    for n, n_spin in ((10, 1), (33, 2), (64, 2)):
        b = scf_b200.distributed.sharded_builder(
            lambda **kw: scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, **kw), dist)
        assert b.shard_range == scf_b200.distributed.shard_range(n, rank, world)
        eri = synth.hash_eri(n)
        da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
        db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
        dens = np.stack([da, db][:n_spin], axis=2)
        total = 2 * da if n_spin == 1 else da + db
        j, k = b.jk(dens, total)
        jr, kr = o.jk_exact(eri, dens, total)
        assert np.abs(j - jr).max() < 1e-12 * np.abs(jr).max()
        assert np.abs(k - kr).max() < 1e-12 * np.abs(kr).max()
        out = np.zeros((n, n, n_spin), order="F")
        b.add_2e_term(out, dens, total)
        ref = o.add_2e_term_exact(np.zeros((n, n, n_spin)), eri, dens, total)
        assert np.abs(out - ref).max() < 1e-12 * np.abs(ref).max()
        # identical on every rank (bitwise: same partials, same NCCL reduction result)
        t = torch.from_numpy(np.ascontiguousarray(out)).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        b.close()
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
