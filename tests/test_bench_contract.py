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
                        "--n-bas", "32", "--cpu-slabs", "4", "--steps", "4", "--warmup", "2"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fock_builds_per_s" and d["unit"] == "Fock builds/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None and d["value"] > 0
    assert d["steps"] == 4 and d["warmup"] == 2                 # the driver's counts are honoured
    # same config keys (and values) as our arm emits for the same arguments
    assert d["config"] == bench.config_dict(32, 2, "full", 1, "p2p")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert "4 of 32 nu-slabs" in cb["sample"] and "scaled x32/4" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--n-bas", "16"], capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_timed_cpu_strategy_is_the_reference_contraction():
    """What the CPU arm times (GEMV for J, blocked permute + GEMV per K line) is the reference's
    contraction: its result equals the oracle's add_2e_term on the sampled nu-slabs."""
    n, s = 12, 5
    ref_arm = bench.CpuReference(n, 2, s)
    out = ref_arm.one_pass()
    full = synth.hash_eri(n)
    da, db = synth.hash_symmetric_matrix(n, synth.SEED_DA), synth.hash_symmetric_matrix(n, synth.SEED_DB)
    ref = o.add_2e_term(np.zeros((n, n, 2)), full, np.stack([da, db], axis=2), da + db)
    for sp in range(2):
        assert np.abs(out[sp] - ref[:, :s, sp]).max() < 1e-12 * np.abs(ref).max()


def test_workload_bytes():
    assert bench.eri_total_bytes(256, "full") == 8 * 256 ** 4
    assert bench.eri_total_bytes(384, "packed8") == 21856961280
