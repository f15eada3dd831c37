"""CPU check of the bench contract's reference arm: `bench.py --impl reference` must print ONE JSON line with the keys the
driver reads (metric/value/unit/config as the product arm, `impl`, `cpu_baseline`, `e2e` with zero copy bytes), on a small
bounded sample so that it runs in seconds here.  The product arm needs a GPU and is exercised by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-points", "6400"] + extra, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    return lines


def test_reference_arm_prints_one_contract_line():
    lines = _run([])
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["higher_is_better"] is True and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["unit"] == "point*comp/s" and j["value"] > 0 and j["n_gpus"] == 1 and j["steps"] == 1
    assert "N=1e+08 d=16 K=64" in j["config"]["workload"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "N_cpu=6400" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    # under torchrun only rank 0 runs and prints; the other ranks exit 0 without work
    assert _run(["--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
