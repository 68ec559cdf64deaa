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
    exchange = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    for n, n_spin, layout in ((10, 1, "full"), (33, 2, "full"), (64, 2, "full"), (40, 2, "packed8"),
                              (7, 1, "full")):
        b = scf_b200.distributed.sharded_builder(
            lambda **kw: scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout=layout, **kw),
            dist, exchange=exchange)
        if layout == "full":
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
        for _ in range(5):      # repeated builds: epochs / double buffering / partial-buffer hygiene
            out2 = np.zeros((n, n, n_spin), order="F")
            b.add_2e_term(out2, dens, total)
            assert np.abs(out2 - ref).max() < 1e-12 * np.abs(ref).max()
        # identical on every rank (full layout: bitwise, same partials and same reduction order)
        t = torch.from_numpy(np.ascontiguousarray(out)).cuda()
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        if layout == "full":
            assert all(torch.equal(g, gathered[0]) for g in gathered)
        b.close()
    # BASELINE-size shard set: n_bf = 256 UHF split over the ranks, exact unit-density answers
    n = 256
    b = scf_b200.distributed.sharded_builder(
        lambda **kw: scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, **kw), dist,
        exchange=exchange)
    idx = np.arange(n)
    unit = scf_b200.synthetic.unit_symmetric
    j, k = b.jk(np.stack([unit(n, 17, 5), unit(n, 255, 255)], axis=2), unit(n, 3, 250))
    e = synth.hash_eri_elements
    assert np.array_equal(j, 2 * e(3, 250, idx[:, None], idx[None, :]))
    assert np.array_equal(k[:, :, 0], e(idx[:, None], 5, 17, idx[None, :]) + e(idx[:, None], 17, 5, idx[None, :]))
    assert np.array_equal(k[:, :, 1], e(idx[:, None], 255, 255, idx[None, :]))
    # dense densities at the BASELINE size: the summed result on EVERY rank against the long-double
    # literal contraction on a nu-slab sample that touches every rank's columns
    from oracle import c_oracle
    rng = np.random.default_rng(2560)
    dens = np.asfortranarray(rng.standard_normal((n, n, 2)))
    total = np.asfortranarray(rng.standard_normal((n, n)))
    slabs = sorted({scf_b200.distributed.shard_range(n, r, world)[0] for r in range(world)} | {n - 1})
    eri_slabs = np.concatenate([c_oracle.hash_eri(n, synth.SEED_ERI, v, v + 1) for v in slabs], axis=3)
    jr, kr = c_oracle.jk_literal_ld_slabs(np.asfortranarray(eri_slabs), dens, total)
    for _ in range(2):
        j, k = b.jk(dens, total)
        assert np.abs(j[:, slabs] - jr).max() < 1e-12 * np.abs(jr).max()
        assert np.abs(k[:, slabs, :] - kr).max() < 1e-12 * np.abs(kr).max()
        out = np.zeros((n, n, 2), order="F")
        b.add_2e_term(out, dens, total)
        assert np.abs(out[:, slabs, :] - (jr[:, :, None] - kr)).max() < 1e-12 * np.abs(kr).max()
    b.close()
    # packed8 at a size with several i-blocks per rank: dense partials really are summed
    n = 130
    b = scf_b200.distributed.sharded_builder(
        lambda **kw: scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout="packed8", **kw), dist,
        exchange=exchange)
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    dens = np.asfortranarray(np.stack([da, db], axis=2))
    slabs = [0, 64, 129]
    eri_slabs = np.concatenate([c_oracle.hash_eri(n, synth.SEED_ERI, v, v + 1) for v in slabs], axis=3)
    jr, kr = c_oracle.jk_literal_ld_slabs(np.asfortranarray(eri_slabs), dens, da + db)
    j, k = b.jk(dens, da + db)
    assert np.abs(j[:, slabs] - jr).max() < 1e-12 * np.abs(jr).max()
    assert np.abs(k[:, slabs, :] - kr).max() < 1e-12 * np.abs(kr).max()
    b.close()
    dist.barrier()
    if rank == 0:
        print(f"mgpu_check ok: world={world} exchange={exchange}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
