"""N > 1 host logic on the CPU: two gloo ranks exchange the communicator id, take their
nu-slab shards and allreduce the partial J/K — must equal the unsharded oracle build."""
import os
import socket
import subprocess
import sys

import pytest

import scf_b200

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_reproduce_unsharded_build(world):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                   LOCAL_RANK=str(rank), WORLD_SIZE=str(world), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py")],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    for rank, p in enumerate(procs):
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert f"rank {rank} ok" in out


@pytest.mark.parametrize("n,world", [(256, 8), (10, 4), (5, 8), (384, 3)])
def test_shard_ranges_tile_the_slabs(n, world):
    r = [scf_b200.distributed.shard_range(n, g, world) for g in range(world)]
    assert r[0][0] == 0 and r[-1][1] == n
    assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in r]
    assert max(sizes) - min(sizes) <= 1
