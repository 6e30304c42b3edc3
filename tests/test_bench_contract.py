"""bench.py's reference arm (the CPU leg, `--impl reference`): one JSON line with the contract's keys on rank 0, nothing on the
other ranks.  Runs the oracle on a 64 x 64 sample, so it needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "64", "--steps", "1",
                           "--warmup", "1", "--gpus", extra_env.get("WORLD_SIZE", "1")],
                          capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line_on_rank_0():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pnp_admm_image_iters_per_sec_256" and d["unit"] == "image-iters/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["dtype"] == "f32"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "sample" in d["cpu_baseline"] and d["config"]["workload"].startswith("batch 64 per GPU of 64x64")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""
