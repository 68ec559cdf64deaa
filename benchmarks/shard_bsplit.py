"""Time one rank's shard of the full-layout Fock build on a single GPU (no exchange), e.g. the
1/8 shard of n_bf=256 that each GPU of an 8-GPU run streams.  Used to tune the b-split rule
(SCFB_FULL_BSPLIT / SCFB_FULL_TAIL override csrc/jk_full.cu:scfb_full_plan; one process per setting because the
override is read once).   python benchmarks/shard_bsplit.py --n-bas 256 --ranks 8 --rank 5"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(a):
    import numpy as np
    import torch
    import scf_b200
    from scf_b200 import synthetic
    n, n_spin = a.n_bas, a.n_spin
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synthetic.SEED_ERI, layout=a.layout, rank=a.rank, n_ranks=a.ranks,
                                         allow_partial=True)
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.stack([synthetic.hash_density(n, synthetic.SEED_DA + s)
                                   for s in range(n_spin)], axis=0)).to(dev).contiguous()
    tot = (d.sum(0) if n_spin == 2 else 2 * d[0]).contiguous()
    out = torch.zeros_like(d)
    b.set_stream(torch.cuda.current_stream().cuda_stream)
    for _ in range(5):
        b.add_2e_term_device(out.data_ptr(), d.data_ptr(), tot.data_ptr(), n_spin)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        b.add_2e_term_device(out.data_ptr(), d.data_ptr(), tot.data_ptr(), n_spin)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(f"bsplit={os.environ.get('SCFB_FULL_BSPLIT', 'auto'):>4}  n={n} rank {a.rank}/{a.ranks} "
          f"n_spin={n_spin}: {ms:.4f} ms/build  {b.eri_bytes / ms / 1e6:.0f} GB/s", flush=True)
    b.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-bas", type=int, default=256)
    ap.add_argument("--n-spin", type=int, default=2)
    ap.add_argument("--ranks", type=int, default=8)
    ap.add_argument("--rank", type=int, default=5)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--layout", default="full")
    ap.add_argument("--sweep", default="auto,1,2,4,8")
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        one(a)
    else:
        for v in a.sweep.split(","):
            env = dict(os.environ)
            env.pop("SCFB_FULL_BSPLIT", None)
            if v != "auto":
                env["SCFB_FULL_BSPLIT"] = v
            subprocess.run([sys.executable, __file__, "--child", "--n-bas", str(a.n_bas), "--n-spin",
                            str(a.n_spin), "--ranks", str(a.ranks), "--rank", str(a.rank), "--steps",
                            str(a.steps), "--layout", a.layout], env=env, check=True)
