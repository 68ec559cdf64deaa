"""CPU checks of bench.py: the reference arm's JSON line follows the contract, and the timed
CPU strategy computes the right thing (its result equals the oracle on the same slabs)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import scf_oracle as o  # noqa: E402
from oracle import synth  # noqa: E402


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--n-bas", "32", "--cpu-slabs", "4", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fock_builds_per_s" and d["unit"] == "Fock builds/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--n-bas", "16"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_timed_cpu_strategy_is_the_reference_contraction():
    n, s = 12, 5
    eri = synth.hash_eri(n, nu_lo=2, nu_hi=2 + s)
    da, db = synth.hash_symmetric_matrix(n, 2), synth.hash_symmetric_matrix(n, 3)
    out = bench.reference_strategy_slabs(eri, [da, db], da + db)
    full = synth.hash_eri(n)
    ref = o.add_2e_term(np.zeros((n, n, 2)), full, np.stack([da, db], axis=2), da + db)
    for sp in range(2):
        assert np.abs(out[sp] - ref[:, 2:2 + s, sp]).max() < 1e-12 * np.abs(ref).max()


def test_workload_bytes():
    class A:
        layout, n_bas, n_spin = "full", 256, 2
    assert bench.eri_total_bytes(A) == 8 * 256 ** 4
    A.layout, A.n_bas = "packed8", 384
    assert bench.eri_total_bytes(A) == 21856961280
