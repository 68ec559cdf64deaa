#!/usr/bin/env python
"""Fock builds/s (J+K) of the B200-native path on BASELINE.json's headline config:
synthetic UHF n_bf=256, full fp64 ERI (34.36 GB) resident in HBM, alpha/beta J+K from one
ERI pass.  One "step" = one `add_2e_term!`-equivalent call (scfb_fock_jk*), the
cross-GPU allreduce included when N > 1.

  python bench.py [--gpus N --steps K --warmup W]          ours (under torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's CPU strategy
                                                           (oracle port; rank 0 only)
Prints ONE JSON line (rank 0).  The inputs (34 GB) are far larger than the 126 MB L2, so
no explicit L2 flush is needed between timed iterations (config.l2_policy says so).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # the CPU arm may use every host thread; torchrun exports OMP_NUM_THREADS=1 to its workers
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.pop(_v, None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "fock_builds_per_s"
UNIT = "Fock builds/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--n-bas", type=int, default=256)
    p.add_argument("--n-spin", type=int, default=2, choices=[1, 2])
    p.add_argument("--layout", default="full", choices=["full", "packed8"])
    p.add_argument("--cpu-slabs", type=int, default=4, help="nu-slabs in the CPU baseline sample")
    p.add_argument("--cpu-steps", type=int, default=3)
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    p.add_argument("--exchange", default="nccl", choices=["nccl", "p2p"],
                   help="cross-GPU sum of the partial J/K: NCCL allreduce or peer-memory stores")
    return p.parse_args()


def eri_total_bytes(a):
    if a.layout == "full":
        return 8 * a.n_bas ** 4
    p = a.n_bas * (a.n_bas + 1) // 2
    return 8 * (p * (p + 1) // 2)


def workload_name(a):
    spin = "UHF" if a.n_spin == 2 else "RHF"
    return (f"synthetic {spin} n_bf={a.n_bas}, {a.layout} fp64 ERI "
            f"({eri_total_bytes(a) / 1e9:.2f} GB), hash family")


def measured_traffic(a, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed `ncu --set full` capture of this exact workload (profiles/r01_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path) or world != 1:
        return None
    key = f"{a.layout}_n{a.n_bas}_spin{a.n_spin}"
    if a.layout == "full" and dominant_kernel(a) == "jk_full_stream_kernel":
        key += "_ldg"
    return json.load(open(path)).get(key, {}).get("dram_bytes_per_launch")


def dominant_kernel(a):
    """Name of the kernel the roofline object describes (same selection rule as csrc/jk_full.cu)."""
    if a.layout != "full":
        return "jk_packed_kernel"
    tma = os.environ.get("SCFB_FULL_TMA", "1") != "0" and a.n_bas % 2 == 0 and 8 <= a.n_bas <= 512
    return "jk_full_tma_kernel" if tma else "jk_full_stream_kernel"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---- nvidia-smi clock sampler -------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for t, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk, mx = float(f[0]), float(f[1])
            except ValueError:
                continue
            if t0 - 0.05 <= t <= t1 + 0.05:
                sm.append(clk)
                try:
                    power.append(float(f[2]))
                except ValueError:
                    pass
                for name, v in zip(self.NAMES, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            smax.append(mx)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- the reference's CPU strategy on a bounded sample (oracle port) ------------------------
def reference_strategy_slabs(eri, dens, total):
    """J - K_s for the nu-slabs held in `eri` ((n,n,n,S), F-ordered), executed the way the
    reference's two @tensor lines execute on a CPU: J = GEMV on the no-copy (ab)x(mu nu) matrix
    view; every K line = permuted copy P[a,b,mu,nu] = eri[mu,b,a,nu] + GEMV."""
    n, n_slabs = eri.shape[0], eri.shape[3]
    n2 = n * n
    m = eri.reshape(n2, n * n_slabs, order="F")
    j = m.T @ total.reshape(n2, order="F")
    out = []
    for d in dens:                                                   # one permuted copy per K line
        p = np.asfortranarray(eri.transpose(2, 1, 0, 3)).reshape(n2, n * n_slabs, order="F")
        out.append((j - p.T @ d.reshape(n2, order="F")).reshape(n, n_slabs, order="F"))
    return out


def cpu_reference_sample(n, n_spin, n_slabs, steps):
    """J as a GEMV over the no-copy (ab)x(mu nu) view, K as permute-copy + GEMV per spin —
    how the reference's @tensor lines execute on the CPU — on `n_slabs`
    nu-slabs of the same synthetic tensor.  builds/s = 1 / (t_sample * n / n_slabs)."""
    from oracle import synth
    n_slabs = min(n_slabs, n)
    eri = synth.hash_eri(n, nu_lo=0, nu_hi=n_slabs)                  # (n,n,n,S) F-ordered
    da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
    db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
    dens = [da, db][:n_spin]
    total = 2 * da if n_spin == 1 else da + db
    n2 = n * n
    times = []
    for _ in range(steps + 1):
        t0 = time.perf_counter()
        reference_strategy_slabs(eri, dens, total)
        times.append(time.perf_counter() - t0)
    t = float(np.median(times[1:]))                                   # first pass = warm-up
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count()
    # extra reference point (clearly NOT what the reference does): a single-pass OpenMP C kernel
    onepass = None
    try:
        from oracle import c_oracle
        density = np.asfortranarray(np.stack(dens, axis=2))
        tt = []
        for _ in range(steps + 1):
            t0 = time.perf_counter()
            c_oracle.jk_onepass(eri, density, total, 0, n_slabs)
            tt.append(time.perf_counter() - t0)
        onepass = 1.0 / (float(np.median(tt[1:])) * n / n_slabs)
    except Exception:                                                 # pragma: no cover
        pass
    return {"value": 1.0 / (t * n / n_slabs), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_slabs} of {n} nu-slabs of the same n_bf={n} hash tensor "
                      f"({8 * n ** 3 * n_slabs / 1e9:.2f} GB), NumPy/OpenBLAS GEMV for J + "
                      f"permute-copy+GEMV per spin for K, median of {steps} passes, scaled x{n}/{n_slabs}",
            "sample_seconds": t, "onepass_openmp_c_value": onepass}


def run_reference(a, rank):
    if rank != 0:
        return
    steps = max(1, min(a.steps, a.cpu_steps if a.cpu_steps else a.steps))
    r = cpu_reference_sample(a.n_bas, a.n_spin, a.cpu_slabs, steps)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT,
            "n_gpus": a.gpus, "steps": steps, "warmup": 1,
            "ms_per_step": 1000.0 / r["value"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "n_bas": a.n_bas, "n_spin": a.n_spin},
            "cpu_baseline": r,
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def verify_known_answer(builder, n, n_spin):
    """Cheap exactness check of the timed builder (every rank): symmetric unit densities pick
    single tensor elements of the hash family, whose values are O(1) closed forms."""
    from scf_b200 import synthetic
    a0, b0, a1, b1 = 3 % n, (n - 5) % n, (n // 2) % n, 1 % n
    unit = lambda a, b: synthetic.unit_symmetric(n, a, b)
    dens = np.stack([unit(a1, b1)] * n_spin, axis=2)
    j, k = builder.jk(dens, unit(a0, b0))
    idx = np.arange(n)
    e = lambda a, b, c, d: synthetic.hash_eri_elements(a, b, c, d)
    w0 = 1.0 if a0 == b0 else 2.0        # E_ab + E_ba picks eri[a,b,..] + eri[b,a,..]
    jr = w0 * e(a0, b0, idx[:, None], idx[None, :])
    kr = e(idx[:, None], b1, a1, idx[None, :]) + (0 if a1 == b1 else e(idx[:, None], a1, b1, idx[None, :]))
    ok = np.abs(j - jr).max() <= 1e-13 * np.abs(jr).max() and np.abs(k[:, :, 0] - kr).max() <= 1e-13 * np.abs(kr).max()
    if not ok:
        raise SystemExit("bench.py: verification of the timed builder FAILED")
    return "unit-density known answers match the closed-form tensor elements"


# ---- ours ---------------------------------------------------------------------------------
def run_ours(a, rank, local_rank, world):
    import torch
    import scf_b200
    from scf_b200 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    comm_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [scf_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm_id = box[0]

    n, n_spin = a.n_bas, a.n_spin
    n2 = n * n
    builder = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synthetic.SEED_ERI, layout=a.layout,
                                               device=local_rank, rank=rank, n_ranks=world)
    if world > 1 and a.exchange == "p2p":
        builder.peer_import(scf_b200.distributed.exchange_peer_handles(dist, builder))
        dist.barrier()
    elif world > 1:
        builder.comm_init(comm_id)
    stream = torch.cuda.Stream()
    builder.set_stream(stream.cuda_stream)

    da = synthetic.hash_density(n, synthetic.SEED_DA)
    db = synthetic.hash_density(n, synthetic.SEED_DB)
    dens_h = np.asfortranarray(np.stack([da, db][:n_spin], axis=2))
    total_h = np.asfortranarray(2 * da if n_spin == 1 else da + db)
    # torch views of the column-major arrays: tensor[s, v, m] == density[m, v, s]
    d_dens = torch.from_numpy(dens_h.T.copy()).cuda()
    d_total = torch.from_numpy(total_h.T.copy()).cuda()
    d_out = torch.zeros_like(d_dens)

    def step():
        builder.add_2e_term_device(d_out.data_ptr(), d_dens.data_ptr(), d_total.data_ptr(), n_spin)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = builder.launch_count
    with torch.cuda.stream(stream):
        for _ in range(a.warmup):
            step()
        barrier()
        launches_w = builder.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_load0 = time.perf_counter()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        barrier()
        t_load1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        launches = builder.launch_count - launches_w

        # roofline pass: device time of the ERI-streaming kernel alone, per launch
        stream_ms = []
        for _ in range(a.steps):
            step()
            stream_ms.append(builder.last_build_ms()[0])
        barrier()

        # end to end through the public host-pointer API (pinned staging inside the library)
        e2e_steps = a.e2e_steps or a.steps

        def pinned_like(arr):    # page-locked host copy with the same column-major layout
            t = torch.empty(arr.T.shape, dtype=torch.float64, pin_memory=True)
            view = t.numpy().T
            view[...] = arr
            return t, view

        _keep_d, dens_p = pinned_like(dens_h)
        _keep_t, total_p = pinned_like(total_h)
        _keep_o, out_h = pinned_like(np.zeros((n, n, n_spin), order="F"))
        assert out_h.flags.f_contiguous and dens_p.flags.f_contiguous
        dens_h, total_h = dens_p, total_p
        for _ in range(min(3, a.warmup)):
            builder.add_2e_term(out_h, dens_h, total_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            builder.add_2e_term(out_h, dens_h, total_h)
        barrier()
        e2e_s = time.perf_counter() - t0
        t_load2 = time.perf_counter()
    clocks = sampler.stop(t_load0, t_load2) if sampler else None
    verified = verify_known_answer(builder, n, n_spin)

    if dist is not None:
        t = torch.tensor([ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        # algorithmic bytes of one stream launch = this rank's share of the UNPADDED tensor
        launch_bytes = float(builder.eri_bytes)
        if a.layout == "packed8":
            launch_bytes *= eri_total_bytes(a) / (8.0 * synthetic.packed8_total(n))
        k_ms = float(np.mean(stream_ms))
        achieved = launch_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": a.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "n_bas": n, "n_spin": n_spin,
                       "layout": a.layout, "sharding": f"{'nu-slabs' if a.layout == 'full' else 'i-blocks'} over {world} rank(s)",
                       "exchange": a.exchange if world > 1 else "none",
                       "l2_policy": "inputs (ERI shard) far larger than the 126 MB L2; no flush"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(a, world),
                         "kernel": dominant_kernel(a),
                         "kernel_ms": k_ms,
                         "algorithmic_bytes_per_launch": launch_bytes, "peak_source": peak_src,
                         "kernel_share_of_step": k_ms / (ms / a.steps)},
            "e2e": {"value": e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": 8 * n2 * (2 * n_spin + 1),
                    "d2h_bytes_per_step": 8 * n2 * n_spin, "steps": e2e_steps},
            "gpu_launches": int(launches), "verified": verified,
            "clocks": clocks,
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference_sample(n, n_spin, a.cpu_slabs, a.cpu_steps)
        print(json.dumps(line), flush=True)
    builder.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank)
        return
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun --nproc-per-node {a.gpus}")
    run_ours(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
