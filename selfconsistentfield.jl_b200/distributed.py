"""One process per GPU: rank g owns a contiguous block of nu-slabs of the
ERI (`shard_range`, same rule as `scfb_split_range` in csrc/scfb_internal.h); per Fock
build the partial [J | K_a | K_b] buffers are summed with ONE ncclAllReduce inside
libscf_b200 (csrc/comm.cu).  torch.distributed is only the plumbing that carries the NCCL
unique id from rank 0 to the other ranks (any backend: nccl on the GPU box, gloo in the
CPU tests)."""
from __future__ import annotations

import os


def env_ranks():
    """(rank, local_rank, world_size) as torchrun exports them."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_bas: int, rank: int, n_ranks: int):
    """nu-slabs [lo, hi) owned by `rank`; ragged remainders go to the first ranks."""
    base, rem = divmod(n_bas, n_ranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_comm_id(dist, rank: int, make_id=None) -> bytes:
    """Rank 0 creates the 128-byte NCCL unique id, everyone receives it."""
    if make_id is None:
        from .builder import comm_unique_id as make_id
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    cid = box[0]
    if not isinstance(cid, (bytes, bytearray)) or len(cid) != 128:
        raise RuntimeError("communicator id exchange failed")
    return bytes(cid)


def exchange_peer_handles(dist, builder) -> bytes:
    """All-gather the 128-byte CUDA IPC handle records of every rank (rank order)."""
    mine = builder.peer_export()
    box = [None] * dist.get_world_size()
    dist.all_gather_object(box, mine)
    if any(not isinstance(x, (bytes, bytearray)) or len(x) != 128 for x in box):
        raise RuntimeError("peer handle exchange failed")
    return b"".join(bytes(x) for x in box)


def sharded_builder(factory, dist=None, exchange="nccl", **kw):
    """Create this rank's builder with `factory(rank=..., n_ranks=..., device=..., **kw)`
    (e.g. JKBuilderCUDA.synthetic / .from_tensor partial) and set up the cross-GPU exchange:
    "nccl" (one ncclAllReduce per build, the default) or "p2p" (peer-memory stores + local sum)."""
    rank, local_rank, world = env_ranks()
    b = factory(rank=rank, n_ranks=world, device=local_rank, **kw)
    if world > 1:
        if dist is None:
            raise ValueError("torch.distributed module required for world_size > 1")
        if exchange == "p2p":
            b.peer_import(exchange_peer_handles(dist, b))
            dist.barrier()          # nobody stores into a peer before everyone has mapped it
        elif exchange == "nccl":
            b.comm_init(exchange_comm_id(dist, rank))
        else:
            raise ValueError(f"unknown exchange {exchange!r}")
    return b
