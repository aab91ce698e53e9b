"""bench.py --impl reference (the CPU arm the driver runs beside ours): one JSON line with the
contract's keys, rank 0 only under torchrun, bounded run time."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, KC_BENCH_REF_SECONDS="3", **env_extra)
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args],
                       capture_output=True, text=True, env=env, timeout=300)
    return r, time.perf_counter() - t0


def test_reference_arm_prints_one_contract_line():
    r, secs = _run({}, "--steps", "3", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["steps"] == 3 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["metric"] == "dwa_trajectory_steps_per_s" and d["unit"] == "trajectory-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert secs < 120


def test_reference_arm_runs_on_rank_zero_only():
    r, _ = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0 and r.stdout.strip() == ""
