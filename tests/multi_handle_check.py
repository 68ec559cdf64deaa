"""Single-process multi-GPU handle at the BASELINE sizes (run on a multi-GPU box):
    python tests/multi_handle_check.py [G]
ONE JKBuilderCUDA over devices 0..G-1 (scfb_create_multi) driven by this one thread through the
host-pointer boundary call — what an unmodified run_scf does (/root/reference/src/run_scf.jl:4-65).
Checks dense-density results against the long-double literal contraction on a nu-slab sample that
touches every member's columns, and reports builds/s of the host-pointer call."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import scf_b200  # noqa: E402
from oracle import c_oracle, synth  # noqa: E402


def check(n, n_spin, layout, devices, steps):
    g = len(devices)
    b = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synth.SEED_ERI, layout=layout, devices=devices)
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    dens = np.asfortranarray(np.stack([da, db][:n_spin], axis=2))
    total = np.asfortranarray(2 * da if n_spin == 1 else da + db)
    slabs = sorted({scf_b200.distributed.shard_range(n, r, g)[0] for r in range(g)} | {n - 1})
    if n >= 384:
        slabs = slabs[::max(1, len(slabs) // 3)][:3] + [n - 1]
    eri = np.concatenate([c_oracle.hash_eri(n, synth.SEED_ERI, v, v + 1) for v in slabs], axis=3)
    jr, kr = c_oracle.jk_literal_ld_slabs(np.asfortranarray(eri), dens, total)
    errs = []
    for _ in range(2):
        j, k = b.jk(dens, total)
        errs.append(max(np.abs(j[:, slabs] - jr).max() / np.abs(jr).max(),
                        np.abs(k[:, slabs, :] - kr).max() / np.abs(kr).max()))
        out = np.zeros((n, n, n_spin), order="F")
        b.add_2e_term(out, dens, total)
        errs.append(np.abs(out[:, slabs, :] - (jr[:, :, None] - kr)).max() / np.abs(kr).max())
    assert max(errs) < 1e-12, errs

    def pinned(arr):
        t = torch.empty(arr.T.shape, dtype=torch.float64, pin_memory=True)
        v = t.numpy().T
        v[...] = arr
        return t, v
    _k1, dp = pinned(dens)
    _k2, tp = pinned(total)
    _k3, op = pinned(np.zeros((n, n, n_spin), order="F"))
    for _ in range(3):
        b.add_2e_term(op, dp, tp)
    b.sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        b.add_2e_term(op, dp, tp)
    b.sync()
    dt = (time.perf_counter() - t0) / steps
    stream_ms, total_ms = b.build_ms_history(min(steps, 128))
    b.close()
    return {"n_bas": n, "n_spin": n_spin, "layout": layout, "n_devices": g, "max_rel_err": float(max(errs)),
            "e2e_builds_per_s": 1.0 / dt, "e2e_ms": dt * 1e3,
            "member0_stream_kernel_ms": float(np.mean(stream_ms)), "member0_build_ms": float(np.mean(total_ms))}


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    devices = list(range(g))
    todo = [(256, 2, "full", 100), (384, 1, "packed8", 50)]
    if 8 * 512 ** 4 / g < 150e9:
        todo.append((512, 1, "full", 20))
    for n, ns, lay, steps in todo:
        print(json.dumps(check(n, ns, lay, devices, steps)), flush=True)
    print(f"multi_handle_check ok: {g} devices, one process, one handle")


if __name__ == "__main__":
    main()
