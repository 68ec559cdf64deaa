#!/usr/bin/env python
"""BASELINE.json config 2: synthetic (chain family) RHF, complete SCF loop with EDIIS->cDIIS,
everything O(n^3)+ on the device.  Prints one JSON line with the per-phase wall time.
  python benchmarks/scf_loop.py [--n-bas 128] [--layout full|packed8] [--unrestricted] [--devices G]
--devices G shards the ERI over GPUs 0..G-1 behind ONE builder of this one process
(scfb_create_multi): BASELINE configs 4 / 5 ("chain for one SCF") as a complete
device-resident SCF — dense algebra on device 0, every GPU streaming its shard each Fock build.
(The energy is checked against the CPU oracle by tests/test_scf_gpu.py::test_config2_*, not here:
only tests/, smoke() and bench.py's CPU baseline may touch oracle/.)"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import scf_b200  # noqa: E402
from scf_b200 import scf as S  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--n-bas", type=int, default=128)
    p.add_argument("--layout", default="full")
    p.add_argument("--unrestricted", action="store_true")
    p.add_argument("--devices", type=int, default=0)
    a = p.parse_args()
    n = a.n_bas
    model = scf_b200.synthetic.ChainModel(n)
    t0 = time.perf_counter()
    kw_dev = {"devices": list(range(a.devices))} if a.devices > 1 else {}
    builder = scf_b200.JKBuilderCUDA.synthetic(n, "chain", 0, aux=model.aux, layout=a.layout, **kw_dev)
    builder.sync()
    t_gen = time.perf_counter() - t0
    problem = scf_b200.ScfProblem(model.overlap, model.n_elec(a.unrestricted), n, not a.unrestricted,
                                  {"nuclear_repulsion": model.nuclear_repulsion},
                                  {"kinetic": model.kinetic, "nuclear_attraction": model.nuclear_attraction},
                                  {"coulomb+exchange": builder})
    guess = scf_b200.compute_guess(problem)
    # instrument the device step
    phases = {"__init__": 0.0, "fock": 0.0, "roothaan": 0.0, "diis_push": 0.0, "diis_extrapolate": 0.0, "get": 0.0}
    for name in list(phases):
        orig = getattr(S.DeviceScf, name)

        def timed(self, *args, _orig=orig, _name=name):
            t = time.perf_counter()
            out = _orig(self, *args)
            phases[_name] += time.perf_counter() - t
            return out
        setattr(S.DeviceScf, name, timed)
    kw = dict(max_error_norm=1e-9, max_energy_total_change=1e-11, max_iter=300)
    scf_b200.run_scf(problem, guess, verbose=False, **kw)          # warm-up (cuSOLVER/cuBLAS init)
    for k in phases:
        phases[k] = 0.0
    t0 = time.perf_counter()
    res = scf_b200.run_scf(problem, guess, verbose=False, **kw)
    t_scf = time.perf_counter() - t0
    line = {"config": f"chain {'UHF' if a.unrestricted else 'RHF'} n_bf={n} {a.layout}, device-resident SCF"
                      + (f", one handle over {a.devices} GPUs" if a.devices > 1 else ""),
            "energy_per_site": res["energies_newest"]["total"] / n,
            "converged": bool(res["converged"]), "n_iter": res["n_iter"],
            "e_total": res["energies_newest"]["total"], "scf_seconds": t_scf,
            "ms_per_iteration": 1e3 * t_scf / res["n_iter"],
            "phase_ms_per_iteration": {k: 1e3 * v / res["n_iter"] for k, v in phases.items()
                                       if k not in ("__init__", "get")},
            "setup_ms": 1e3 * phases["__init__"], "result_download_ms": 1e3 * phases["get"],
            "eri_generate_seconds": t_gen, "fock_kernel_ms": builder.last_build_ms()[0]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
