"""Work-item plan of the full-layout kernel (csrc/jk_full.cu: scfb_full_plan, read through
scfb_full_plan_query — a host function, no GPU): the persistent kernel's item decode, restated here,
must tile every (nu-slab, c-chunk, b) of a shard exactly once, address distinct partial rows inside
the allocation, and bring every completion counter to the value its finisher waits for."""
import ctypes as C

import pytest

from scf_b200 import capi

TC = 8


def plan(n, n_slabs, n_sm=148):
    out = (C.c_int * 6)()
    assert capi.lib().scfb_full_plan_query(n, n_slabs, n_sm, out) == 0
    return dict(zip(("tma", "whole_slabs", "bs", "kpart_rows", "jpart_rows", "n_groups"), out))


def decode(item, n, n_chunks, whole_slabs, bs):
    """jk_full_tma_kernel: item -> (slab, chunk, b_lo, b_hi, split)."""
    n_whole = whole_slabs * n_chunks
    if item < n_whole:
        return item // n_chunks, item % n_chunks, 0, n, False
    j = item - n_whole
    bz, cs = j % bs, j // bs
    nb = n // bs
    return whole_slabs + cs // n_chunks, cs % n_chunks, bz * nb, bz * nb + nb, True


@pytest.mark.parametrize("n,n_slabs", [(256, 256), (256, 32), (256, 5), (128, 128), (512, 64), (384, 48), (320, 7),
                                       (64, 8), (8, 1), (130, 130)])
def test_items_tile_the_shard_once(n, n_slabs):
    p = plan(n, n_slabs)
    assert p["tma"] == 1
    n_chunks = (n + TC - 1) // TC
    split_slabs = n_slabs - p["whole_slabs"]
    assert 0 <= split_slabs <= n_slabs and (p["bs"] == 1) == (split_slabs == 0)
    assert n % p["bs"] == 0
    n_items = p["whole_slabs"] * n_chunks + split_slabs * n_chunks * p["bs"]
    assert p["kpart_rows"] == n_items and p["jpart_rows"] == split_slabs * n_chunks * p["bs"]
    assert p["n_groups"] == split_slabs * n_chunks
    covered = {}
    slab_done = [0] * n_slabs
    grp_done = [0] * p["n_groups"]
    for item in range(n_items):
        slab, chunk, b_lo, b_hi, split = decode(item, n, n_chunks, p["whole_slabs"], p["bs"])
        assert 0 <= slab < n_slabs and 0 <= chunk < n_chunks and 0 <= b_lo < b_hi <= n
        for b in range(b_lo, b_hi):
            key = (slab, chunk, b)
            assert key not in covered
            covered[key] = item
        if split:
            grp = (item - p["whole_slabs"] * n_chunks) // p["bs"]
            assert grp == (slab - p["whole_slabs"]) * n_chunks + chunk    # the group a piece folds into
            grp_done[grp] += 1
            if grp_done[grp] == p["bs"]:
                slab_done[slab] += 1
        else:
            slab_done[slab] += 1
    assert len(covered) == n_slabs * n_chunks * n
    assert all(c == p["bs"] for c in grp_done) and all(c == n_chunks for c in slab_done)
    # the tail is finer, never coarser: split slabs are the LAST ones of the counter
    if split_slabs:
        assert p["bs"] in (2, 4, 8) and TC * n * n * 8 // p["bs"] >= 1 << 20


def test_overrides(monkeypatch):
    monkeypatch.setenv("SCFB_FULL_TAIL", "0")
    p = plan(256, 32)
    assert p["bs"] == 1 and p["whole_slabs"] == 32 and p["jpart_rows"] == 0
    monkeypatch.setenv("SCFB_FULL_TAIL", "3")
    assert plan(256, 32)["whole_slabs"] == 29
    monkeypatch.delenv("SCFB_FULL_TAIL")
    monkeypatch.setenv("SCFB_FULL_BSPLIT", "2")
    p = plan(64, 8)
    assert p["bs"] == 2 and p["whole_slabs"] == 0 and p["kpart_rows"] == 8 * 8 * 2
    monkeypatch.setenv("SCFB_FULL_BSPLIT", "3")       # does not divide n: ignored
    assert plan(64, 8)["bs"] == 1
    monkeypatch.delenv("SCFB_FULL_BSPLIT")
    monkeypatch.setenv("SCFB_FULL_TMA", "0")           # the LDG kernel's rule: every item split, k_reduce_kernel sums
    p = plan(256, 32)
    assert p["tma"] == 0 and p["bs"] in (1, 2, 4) and p["kpart_rows"] == 32 * 32 * p["bs"]


def test_odd_sizes_use_the_ldg_kernel():
    p = plan(255, 32)
    assert p["tma"] == 0
