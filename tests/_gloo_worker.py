"""Worker for tests/test_distributed_gloo.py (world_size-2 gloo ranks on the CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import scf_b200  # noqa: E402
from oracle import scf_oracle as o  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, _, world = scf_b200.distributed.env_ranks()
    assert world == dist.get_world_size() and rank == dist.get_rank()
    # 1. communicator-id exchange (fake id: libnccl needs no part in the plumbing test)
    cid = scf_b200.distributed.exchange_comm_id(dist, rank, make_id=lambda: bytes(range(128)))
    assert cid == bytes(range(128))
    # 2. this rank's shard rule matches the oracle's statement of the partition
    n, n_spin = 11, 2
    lo, hi = scf_b200.distributed.shard_range(n, rank, world)
    assert (lo, hi) == o.shard_range(n, rank, world)
    # 3. sum over ranks of the per-shard partial [J|Ka|Kb] == unsharded build
    rng = np.random.default_rng(5)
    eri = rng.standard_normal((n, n, n, n))
    dens = rng.standard_normal((n, n, n_spin))
    total = dens[:, :, 0] + dens[:, :, 1]
    part = np.zeros((1 + n_spin, n, n))
    sub = eri[:, :, :, lo:hi]
    part[0][:, lo:hi] = np.einsum("abmn,ab->mn", sub, total)
    for s in range(n_spin):
        part[1 + s][:, lo:hi] = np.einsum("mban,ab->mn", sub, dens[:, :, s])
    t = torch.from_numpy(part)
    dist.all_reduce(t)                      # the one collective of a Fock build
    j = o.coulomb(eri, total)
    assert np.abs(t[0].numpy() - j).max() < 1e-12 * np.abs(j).max()
    for s in range(n_spin):
        k = o.exchange(eri, dens[:, :, s])
        assert np.abs(t[1 + s].numpy() - k).max() < 1e-12 * np.abs(k).max()
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
