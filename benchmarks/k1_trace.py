"""Timeline of the persistent full-layout kernel (K1) on one shard: where a build's time goes
beyond bytes / bandwidth.  Needs the tuning build of the library with the probes compiled in:

  make -C selfconsistentfield.jl_b200/csrc trace
  python benchmarks/k1_trace.py --n-bas 256 --ranks 8 --rank 5

Per CTA the kernel records (globaltimer, ns): start, first stage landed, item loop left, end; the
number of items, the time inside item epilogues and slab finishers, and the start of its last item.
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SCFB_LIB", os.path.join(ROOT, "selfconsistentfield.jl_b200", "libscf_b200_trace.so"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-bas", type=int, default=256)
    ap.add_argument("--n-spin", type=int, default=2)
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--rank", type=int, default=5)
    ap.add_argument("--builds", type=int, default=5)
    a = ap.parse_args()

    import numpy as np
    import torch
    import scf_b200
    from scf_b200 import synthetic

    lib = ctypes.CDLL(os.environ["SCFB_LIB"])
    lib.scfb_debug_k1_trace.argtypes = [ctypes.c_void_p]
    n, n_spin = a.n_bas, a.n_spin
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synthetic.SEED_ERI, layout="full", rank=a.rank,
                                         n_ranks=a.ranks, allow_partial=True)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.stack([synthetic.hash_density(n, synthetic.SEED_DA + s)
                                   for s in range(n_spin)], axis=0)).to(dev).contiguous()
    tot = (d.sum(0) if n_spin == 2 else 2 * d[0]).contiguous()
    out = torch.zeros_like(d)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(20):
        b.add_2e_term_device(out.data_ptr(), d.data_ptr(), tot.data_ptr(), n_spin)
    torch.cuda.synchronize()
    trace = np.zeros((1024, 8), dtype=np.uint64)
    for it in range(a.builds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # back-to-back pair: the second build is the one a timed loop sees (no idle gap before it)
        b.add_2e_term_device(out.data_ptr(), d.data_ptr(), tot.data_ptr(), n_spin)
        e0.record()
        b.add_2e_term_device(out.data_ptr(), d.data_ptr(), tot.data_ptr(), n_spin)
        e1.record()
        torch.cuda.synchronize()
        rc = lib.scfb_debug_k1_trace(trace.ctypes.data)
        assert rc == 0, rc
        t = trace.astype(np.int64)
        live = t[:, 0] > 0
        t = t[live]
        g = len(t)
        t0 = t[:, 0].min()
        us = lambda x: (x - t0) / 1e3
        start, first, loop_end, end = us(t[:, 0]), us(t[:, 1]), us(t[:, 2]), us(t[:, 3])
        items = t[:, 4] & 0xffffffff
        smid = t[:, 4] >> 32
        epi, fin = t[:, 5] / 1e3, t[:, 7] / 1e3
        last_start = us(t[:, 6])
        T = end.max()
        q = lambda v: "min %.1f  p10 %.1f  med %.1f  p90 %.1f  max %.1f" % (
            v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max())
        bytes_ = b.eri_bytes
        print(f"--- build {it}: event {e0.elapsed_time(e1) * 1e3:.1f} us, first CTA start -> last CTA end {T:.1f} us, "
              f"{g} CTAs on {len(set(smid.tolist()))} SMs, {int(items.sum())} items, "
              f"bytes/T = {bytes_ / T / 1e3:.0f} GB/s")
        print("  CTA start          ", q(start))
        print("  first stage landed ", q(first))
        print("  (landed - start)   ", q(first - start))
        print("  last item start    ", q(last_start))
        print("  item loop left     ", q(loop_end))
        print("  T - loop left      ", q(T - loop_end))
        print("  items per CTA      ", np.bincount(items.astype(int)).tolist())
        print("  epilogues per CTA  ", q(epi), " per item %.2f us" % (epi.sum() / items.sum()))
        print("  slab finishers/CTA ", q(fin))
        # how many CTAs are still streaming at T - x
        for x in (5, 10, 20, 40, 80, 120, 160):
            print(f"    streaming at T-{x:3d} us: {(loop_end > T - x).sum():4d} CTAs on "
                  f"{len(set(smid[loop_end > T - x].tolist())):3d} SMs")
        # mean item time of the CTAs by item count
        dur = (loop_end - first)
        for c in sorted(set(items.tolist())):
            m = items == c
            print(f"    CTAs with {c} items: {m.sum():4d}, stream time {dur[m].mean():.1f} us = {dur[m].mean() / c:.1f} us/item")
    b.close()


if __name__ == "__main__":
    main()
