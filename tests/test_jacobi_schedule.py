"""Host model of the pairing schedules of csrc/eig_jacobi.cu (no GPU): the index arithmetic of the
kernels restated line by line, checked for the invariants the Jacobi iteration relies on —

  * one-CTA kernel: the round-robin tournament pairs every two columns exactly once per sweep and a
    round's pairs are disjoint (they are rotated concurrently);
  * cluster kernels: intra-block tournaments (with a bye for odd block widths) + the block ring
    (L_0 stays, R_0 -> L_1, L_k -> L_k+1, R_k -> R_k-1, L_last -> R_last) make every two of the
    2 * CL * bw columns meet exactly once per sweep, for the double-buffered exchange and for the
    three-buffer / two-phase exchange of the n > 448 variant — and in the latter no block buffer is
    overwritten before its owner has pushed or kept what it held.

The GPU parity of the kernels themselves is tests/test_eig_jacobi_gpu.py."""
import itertools

import pytest


def tournament_pair(m, r, g):
    """jacobi_eig_kernel / the intra-block tournament: pair g of round r among m (even) players."""
    if g == 0:
        return m - 1, r
    p = r + g
    if p >= m - 1:
        p -= m - 1
    q = r - g
    if q < 0:
        q += m - 1
    return p, q


@pytest.mark.parametrize("m", [2, 4, 6, 8, 12, 16, 22, 128, 168])
def test_one_cta_tournament_meets_every_pair_once(m):
    seen = set()
    for r in range(m - 1):
        cols = []
        for g in range(m // 2):
            p, q = tournament_pair(m, r, g)
            assert p != q and 0 <= p < m and 0 <= q < m
            cols += [p, q]
            key = (min(p, q), max(p, q))
            assert key not in seen
            seen.add(key)
        assert len(set(cols)) == m            # disjoint pairs: the round can run concurrently
    assert len(seen) == m * (m - 1) // 2


def sweep_pairs(cl, bw, triple):
    """Column pairs met in one sweep of jacobi_eig_cluster_kernel<cl, ..., triple>; columns are
    (block, j).  Simulates the physical block buffers of every CTA and the pushes."""
    nbuf = 3 if triple else 4
    buf = [[None] * nbuf for _ in range(cl)]          # buf[rank][physical buffer] = block id
    for rank in range(cl):
        buf[rank][0], buf[rank][1] = rank, cl + rank   # load: block rank -> L, block CL + rank -> R
    iL, iR, iS = 0, 1, 2
    met = []
    mt = bw + (bw & 1)

    def resident(rank):
        fixed0 = triple and rank == 0
        return buf[rank][0 if fixed0 else iL], buf[rank][1 if fixed0 else iR]

    # intra-block tournament (both resident blocks of every CTA)
    for rank in range(cl):
        for blk in resident(rank):
            for r in range(mt - 1):
                used = []
                for pi in range(mt // 2):
                    p, q = tournament_pair(mt, r, pi)
                    if p < bw and q < bw:
                        met.append(((blk, p), (blk, q)))
                        used += [p, q]
                assert len(set(used)) == len(used)
    for step in range(2 * cl - 1):
        for rank in range(cl):
            bl, br = resident(rank)
            assert bl is not None and br is not None and bl != br
            for r in range(bw):
                ys = []
                for g in range(bw):
                    q = g + r
                    if q >= bw:
                        q -= bw
                    met.append(((bl, g), (br, q)))
                    ys.append(q)
                assert len(set(ys)) == bw
        if not triple:
            nL, nR = iL ^ 2, iR ^ 2
            new = [row[:] for row in buf]
            for rank in range(cl):
                if rank == 0:
                    new[0][nL] = buf[0][iL]
                    new[1][nL] = buf[0][iR]
                else:
                    last = rank == cl - 1
                    new[rank if last else rank + 1][nR if last else nL] = buf[rank][iL]
                    new[rank - 1][nR] = buf[rank][iR]
            buf = new
            iL, iR = nL, nR
        else:
            # phase 1: L blocks (R of rank 0) into the right neighbour's spare
            live = [[True] * nbuf for _ in range(cl)]    # does the buffer still hold something its owner needs?
            for rank in range(cl):
                live[rank][2 if rank == 0 else iS] = False
            new = [row[:] for row in buf]
            for rank in range(cl):
                if rank == 0:
                    assert not live[1][iS]
                    new[1][iS] = buf[0][1]
                    live[0][1] = False                    # pushed out: rank 0's R buffer is free
                elif rank < cl - 1:
                    assert not live[rank + 1][iS]
                    new[rank + 1][iS] = buf[rank][iL]
                    live[rank][iL] = False
            buf = new
            # phase 2 (after the cluster barrier): R blocks into the left neighbour's vacated buffer
            new = [row[:] for row in buf]
            for rank in range(1, cl):
                dst = (0, 1) if rank == 1 else (rank - 1, iL)
                assert not live[dst[0]][dst[1]], "phase 2 would overwrite a block still in use"
                new[dst[0]][dst[1]] = buf[rank][iR]
            buf = new
            iL, iR, iS = iS, iL, iR
    return met


@pytest.mark.parametrize("cl,bw,triple", [(8, 1, False), (8, 2, False), (8, 5, False), (8, 8, False), (8, 11, False),
                                          (16, 12, False), (16, 14, False), (16, 15, True), (16, 16, True),
                                          (4, 3, True), (4, 2, False)])
def test_cluster_sweep_meets_every_pair_once(cl, bw, triple):
    met = sweep_pairs(cl, bw, triple)
    keys = [tuple(sorted(pair)) for pair in met]
    assert len(keys) == len(set(keys)), "a pair of columns met twice in one sweep"
    cols = [(b, j) for b in range(2 * cl) for j in range(bw)]
    assert set(keys) == set(itertools.combinations(sorted(cols), 2))


@pytest.mark.parametrize("n,cl", [(128, 8), (168, 8), (320, 8), (384, 16), (448, 16), (512, 16)])
def test_shared_memory_budget_of_the_launch_rule(n, cl):
    """scfb_launch_jacobi_eig: block width, thread count and dynamic shared memory of every size class
    stay within a CTA's limits (226 KB opt-in shared memory, the kernels' launch bounds)."""
    ld = n + (n & 1)
    bw = (ld + 2 * cl - 1) // (2 * cl)
    mt = bw + (bw & 1)
    triple = n > 448
    nbuf = 3 if triple else 4
    smem = (nbuf * bw * ld + nbuf * bw + 2 * cl * bw) * 8
    assert smem <= 227 * 1024 - 1024
    assert 32 * max(mt, bw) <= (768 if cl == 8 else 512)
    assert (ld // 2 + 31) // 32 <= (5 if cl == 8 else (8 if triple else 7))
