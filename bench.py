#!/usr/bin/env python
"""Fock builds/s (J+K) of the B200-native path on BASELINE.json's headline config:
synthetic UHF n_bf=256, full fp64 ERI (34.36 GB) resident in HBM, alpha/beta J+K from one
ERI pass.  One "step" = one `add_2e_term!`-equivalent call (scfb_fock_jk*), the cross-GPU
exchange included when N > 1.

  python bench.py [--gpus N --steps K --warmup W]          ours (under torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's CPU strategy
                                                           (oracle port; rank 0 only)
Prints ONE JSON line (rank 0).  Besides the headline the line carries `extra_configs`: BASELINE
config 4 (RHF n_bf=384, packed8) at the same N and, when a shard fits (N >= 4), config 5 (RHF
n_bf=512, full layout), each with its own roofline / e2e / verification.  The inputs (>= 2 GB per
GPU) are far larger than the 126 MB L2, so no explicit L2 flush is needed between timed iterations
(config.l2_policy says so).
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
    p.add_argument("--cpu-slabs", type=int, default=0, help="nu-slabs in the CPU sample (0 = ~1 GB worth)")
    p.add_argument("--cpu-steps", type=int, default=3, help="passes of the cpu_baseline leg of our arm")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-extra", action="store_true", help="headline only (no extra_configs)")
    p.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    p.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                   help="cross-GPU sum of the partial J/K: peer-memory stores fused into the stream kernel "
                        "(default) or one ncclAllReduce per build")
    return p.parse_args()


def eri_total_bytes(n_bas, layout):
    if layout == "full":
        return 8 * n_bas ** 4
    p = n_bas * (n_bas + 1) // 2
    return 8 * (p * (p + 1) // 2)


def workload_name(n_bas, n_spin, layout):
    spin = "UHF" if n_spin == 2 else "RHF"
    return (f"synthetic {spin} n_bf={n_bas}, {layout} fp64 ERI "
            f"({eri_total_bytes(n_bas, layout) / 1e9:.2f} GB), hash family")


def config_dict(n_bas, n_spin, layout, world, exchange):
    """Identical for both arms (`--impl reference` describes its bounded sample in `cpu_baseline`)."""
    return {"workload": workload_name(n_bas, n_spin, layout), "n_bas": n_bas, "n_spin": n_spin,
            "layout": layout,
            "sharding": f"{'nu-slabs' if layout == 'full' else 'i-blocks'} over {world} rank(s)",
            "exchange": exchange if world > 1 else "none",
            "l2_policy": "inputs (ERI shard) far larger than the 126 MB L2; no flush"}


def dominant_kernel(n_bas, layout):
    """Name of the kernel the roofline object describes (same selection rule as csrc/jk_full.cu)."""
    if layout != "full":
        return "jk_packed_kernel"
    tma = os.environ.get("SCFB_FULL_TMA", "1") != "0" and n_bas % 2 == 0 and 8 <= n_bas <= 512
    return "jk_full_tma_kernel" if tma else "jk_full_stream_kernel"


def measured_traffic(n_bas, n_spin, layout, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    committed `ncu --set full` capture of this exact workload (profiles/r02_traffic.json)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(path) or world != 1:
        return None
    key = f"{layout}_n{n_bas}_spin{n_spin}"
    if layout == "full" and dominant_kernel(n_bas, layout) == "jk_full_stream_kernel":
        key += "_ldg"
    return json.load(open(path)).get(key, {}).get("dram_bytes_per_launch")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count()


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
class CpuReference:
    """J - K_s for a SAMPLE of nu-slabs of the same synthetic tensor, executed the way the
    reference's two @tensor lines execute on a CPU: J = GEMV on the no-copy
    (ab) x (mu nu) matrix view; every K line = permuted temporary P[a,b,mu,nu] = eri[mu,b,a,nu]
    (cache-blocked all-core transpose, oracle/jk_oracle.c: permute_mban_slabs) + GEMV, with
    NumPy/OpenBLAS on every host core.  One pass over S of n slabs costs S/n of a build (the work is
    linear in the slab count), so builds/s = 1 / (t_pass * n / S)."""

    def __init__(self, n, n_spin, n_slabs=0, passes=4, budget_s=100.0):
        """n_slabs = 0: the sample is sized so that `passes` passes take about `budget_s` seconds
        (calibrated on a ~1 GB pass), capped by the tensor itself ("full tensor") and by host RAM."""
        from oracle import c_oracle, synth
        self.c_oracle = c_oracle
        self.n, self.n_spin = n, n_spin
        da = synth.hash_symmetric_matrix(n, synth.SEED_DA)
        db = synth.hash_symmetric_matrix(n, synth.SEED_DB)
        self.dens = [da, db][:n_spin]
        self.total = 2 * da if n_spin == 1 else da + db
        auto = n_slabs <= 0
        self._alloc(min(n, n_slabs if not auto else max(1, int(round(1.07e9 / (8.0 * n ** 3))))))
        if auto and self.n_slabs < n:
            self.one_pass()
            t0 = time.perf_counter()
            self.one_pass()
            t1 = time.perf_counter() - t0
            want = int(self.n_slabs * (budget_s / max(passes, 1)) / max(t1, 1e-6))
            try:
                avail = int(next(ln for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")).split()[1]) * 1024
            except (OSError, StopIteration):
                avail = 8 << 30
            want = min(want, int(0.5 * avail / (2 * 8.0 * n ** 3)), n)
            if want > self.n_slabs:
                self._alloc(want)

    def _alloc(self, n_slabs):
        from oracle import synth
        self.n_slabs = n_slabs
        self.eri = self.tmp = None
        self.eri = self.c_oracle.hash_eri(self.n, synth.SEED_ERI, 0, n_slabs)   # (n,n,n,S) F-ordered
        self.tmp = np.empty_like(self.eri, order="F")

    def one_pass(self):
        n, s = self.n, self.n_slabs
        n2 = n * n
        m = self.eri.reshape(n2, n * s, order="F")
        j = m.T @ self.total.reshape(n2, order="F")
        out = []
        for d in self.dens:                                            # one permuted temporary per K line
            p = self.c_oracle.permute_mban(self.eri, self.tmp).reshape(n2, n * s, order="F")
            out.append((j - p.T @ d.reshape(n2, order="F")).reshape(n, s, order="F"))
        return out

    def describe(self, t_pass, passes):
        n, s = self.n, self.n_slabs
        return (f"{'the full tensor: ' if s == n else ''}{s} of {n} nu-slabs of the same n_bf={n} hash tensor "
                f"({8 * n ** 3 * s / 1e9:.2f} GB per pass), "
                f"OpenBLAS GEMV for J + blocked OpenMP permute + GEMV per spin for K, median of {passes} passes "
                f"({t_pass:.3f} s each), scaled x{n}/{s} to one full build (work is linear in the slab count)")


def cpu_baseline_leg(n, n_spin, n_slabs, passes):
    """The cpu_baseline object of our arm's line (rank 0, N = 1)."""
    ref = CpuReference(n, n_spin, n_slabs, passes + 1, 20.0)
    times = []
    for _ in range(passes + 1):
        t0 = time.perf_counter()
        ref.one_pass()
        times.append(time.perf_counter() - t0)
    t = float(np.median(times[1:]))                                    # first pass = warm-up
    onepass = None                  # extra point (NOT what the reference does): single-pass OpenMP C kernel
    try:
        density = np.asfortranarray(np.stack(ref.dens, axis=2))
        tt = []
        for _ in range(passes + 1):
            t0 = time.perf_counter()
            ref.c_oracle.jk_onepass(ref.eri, density, ref.total, 0, ref.n_slabs)
            tt.append(time.perf_counter() - t0)
        onepass = 1.0 / (float(np.median(tt[1:])) * n / ref.n_slabs)
    except Exception:                                                 # pragma: no cover
        pass
    return {"value": 1.0 / (t * n / ref.n_slabs), "unit": UNIT, "cores": host_cores(), "kind": "port",
            "sample": ref.describe(t, passes), "sample_seconds": t, "onepass_openmp_c_value": onepass}


def run_reference(a, rank, world):
    """`--impl reference`: W untimed + K timed steps, each one pass of the reference's CPU strategy
    over the bounded slab sample; value = K steps / their time, scaled to full builds."""
    if rank != 0:
        return
    ref = CpuReference(a.n_bas, a.n_spin, a.cpu_slabs, a.steps + a.warmup, 100.0)
    for _ in range(a.warmup):
        ref.one_pass()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ref.one_pass()
    t = (time.perf_counter() - t0) / max(a.steps, 1)
    value = 1.0 / (t * a.n_bas / ref.n_slabs)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(a.n_bas, a.n_spin, a.layout, world, a.exchange),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": host_cores(), "kind": "port",
                             "sample": ref.describe(t, a.steps), "sample_seconds": t},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- verification of the timed builder -----------------------------------------------------
def verify_known_answer(builder, n, n_spin):
    """Exactness check: symmetric unit densities pick single tensor elements of the hash family,
    whose values are O(1) closed forms (an index check)."""
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
        raise SystemExit("bench.py: unit-density verification of the timed builder FAILED")


def verify_dense(builder, n, n_spin, layout, dens, total, rank, world):
    """Sum check on EVERY rank: the result of the timed builder (after the cross-GPU exchange) for the
    dense benchmark densities against the long-double literal contraction (oracle/jk_oracle.c:
    jk_literal_ld_slabs, JKBuilderFromTensor.jl:36-38,43-47) on a sample of nu-slabs that includes a
    column owned by this rank and one owned by another."""
    from oracle import c_oracle, synth
    lo, hi = (n * rank // world, n * (rank + 1) // world)
    slabs = sorted({min(lo, n - 1), (hi + n // 2) % n, n - 1}) if n >= 384 else \
        sorted({min(lo, n - 1), max(hi - 1, 0), (hi + n // 2) % n, n - 1})
    eri = np.concatenate([c_oracle.hash_eri(n, synth.SEED_ERI, v, v + 1) for v in slabs], axis=3)
    jr, kr = c_oracle.jk_literal_ld_slabs(np.asfortranarray(eri), dens, total)
    j, k = builder.jk(dens, total)
    ej = np.abs(j[:, slabs] - jr).max() / np.abs(jr).max()
    ek = np.abs(k[:, slabs, :] - kr).max() / np.abs(kr).max()
    if not (ej < 1e-12 and ek < 1e-12):
        raise SystemExit(f"bench.py: dense verification of the timed builder FAILED on rank {rank}: {ej:.2e} {ek:.2e}")
    return max(ej, ek), len(slabs)


# ---- ours ---------------------------------------------------------------------------------
def time_config(a, n, n_spin, layout, rank, local_rank, world, dist, comm_id, stream, steps, warmup, sampler_on):
    """Build, time and verify one configuration; returns the JSON fragment (rank 0) or None."""
    import torch
    import scf_b200
    from scf_b200 import synthetic

    n2 = n * n
    builder = scf_b200.JKBuilderCUDA.synthetic(n, "hash", synthetic.SEED_ERI, layout=layout,
                                               device=local_rank, rank=rank, n_ranks=world)
    if world > 1 and a.exchange == "p2p":
        builder.peer_import(scf_b200.distributed.exchange_peer_handles(dist, builder))
        dist.barrier()
    elif world > 1:
        builder.comm_init(comm_id)
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

    sampler = ClockSampler(local_rank) if sampler_on else None
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
        barrier()
        launches_w = builder.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_load0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        t_load1 = time.perf_counter()
        ms = e0.elapsed_time(e1)
        launches = builder.launch_count - launches_w
        # device time of the ERI-streaming kernel alone in THOSE builds (the library's own CUDA
        # events on the same stream, read back after the loop: no synchronisation in between)
        stream_ms, _ = builder.build_ms_history(min(steps, 128))

        # end to end through the public host-pointer API (H2D of density, total density and out from
        # page-locked host arrays and D2H of out inside the call)
        e2e_steps = a.e2e_steps or steps

        def pinned_like(arr):    # page-locked host copy with the same column-major layout
            t = torch.empty(arr.T.shape, dtype=torch.float64, pin_memory=True)
            view = t.numpy().T
            view[...] = arr
            return t, view

        _keep_d, dens_p = pinned_like(dens_h)
        _keep_t, total_p = pinned_like(total_h)
        _keep_o, out_h = pinned_like(np.zeros((n, n, n_spin), order="F"))
        assert out_h.flags.f_contiguous and dens_p.flags.f_contiguous
        for _ in range(min(3, warmup)):
            builder.add_2e_term(out_h, dens_p, total_p)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            builder.add_2e_term(out_h, dens_p, total_p)
        barrier()
        e2e_s = time.perf_counter() - t0
        t_load2 = time.perf_counter()
    clocks = sampler.stop(t_load0, t_load2) if sampler else None
    verify_known_answer(builder, n, n_spin)
    err, n_slabs = verify_dense(builder, n, n_spin, layout, dens_h, total_h, rank, world)
    verified = (f"dense benchmark densities vs the long-double literal contraction on {n_slabs} nu-slabs per rank "
                f"(every rank, after the exchange): max rel. error {err:.1e} < 1e-12; unit-density known answers exact")

    k_ms = float(np.mean(stream_ms)) if len(stream_ms) else float("nan")
    launch_bytes = float(builder.eri_bytes)
    if layout == "packed8":
        launch_bytes *= eri_total_bytes(n, layout) / (8.0 * synthetic.packed8_total(n))
    if dist is not None:
        t = torch.tensor([ms, e2e_s, err], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, err = t.tolist()
    builder.close()
    del d_dens, d_total, d_out
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    peak, peak_src = peaks()
    achieved = launch_bytes / (k_ms * 1e-3) / 1e9
    return {
        "value": steps / (ms * 1e-3), "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "config": config_dict(n, n_spin, layout, world, a.exchange),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": measured_traffic(n, n_spin, layout, world),
                     "kernel": dominant_kernel(n, layout), "kernel_ms": k_ms,
                     "algorithmic_bytes_per_launch": launch_bytes, "peak_source": peak_src,
                     "kernel_share_of_step": k_ms / (ms / steps),
                     "timing": "library CUDA events around the stream kernel inside the timed loop"},
        "e2e": {"value": e2e_steps / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": 8 * n2 * (2 * n_spin + 1),
                "d2h_bytes_per_step": 8 * n2 * n_spin, "steps": e2e_steps},
        "gpu_launches": int(launches), "verified": verified, "clocks": clocks,
    }


def run_ours(a, rank, local_rank, world):
    import torch
    import scf_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    comm_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.Stream()

    def new_comm_id():
        if world == 1 or a.exchange != "nccl":
            return None
        box = [scf_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    head = time_config(a, a.n_bas, a.n_spin, a.layout, rank, local_rank, world, dist, new_comm_id(), stream,
                       a.steps, a.warmup, rank == 0)
    extras = []
    if not a.no_extra and (a.n_bas, a.n_spin, a.layout) == (256, 2, "full"):
        todo = [(384, 1, "packed8")]
        free, _ = torch.cuda.mem_get_info()
        if 8 * 512 ** 4 / world < 0.85 * free:       # BASELINE config 5 needs >= 4 GPUs of 180 GB
            todo.append((512, 1, "full"))
        for n, ns, lay in todo:
            steps = min(a.steps, 30)
            r = time_config(a, n, ns, lay, rank, local_rank, world, dist, new_comm_id(), stream, steps,
                            min(a.warmup, 5), False)
            if r is not None:
                r.pop("clocks")
                r.update({"metric": METRIC, "unit": UNIT, "n_gpus": world})
                extras.append(r)

    if rank == 0:
        line = {"metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": head["config"], "roofline": head["roofline"],
                "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "verified": head["verified"],
                "clocks": head["clocks"]}
        if extras:
            line["extra_configs"] = extras
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_leg(a.n_bas, a.n_spin, a.cpu_slabs, a.cpu_steps)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.impl == "reference":
        run_reference(a, rank, world)
        return
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit(f"--gpus {a.gpus} needs torchrun --nproc-per-node {a.gpus}")
    run_ours(a, rank, local_rank, world)


if __name__ == "__main__":
    main()
