"""bench.py contract, CPU side: the reference arm (--impl reference) runs without a GPU and prints one JSON line
with the keys the driver reads; non-zero ranks of a multi-rank reference launch do no work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gaussians", "2000",
                          "--cpu-sample-ellipsoids", "60", "--steps", "2", "--warmup", "1"], capture_output=True, text=True,
                         env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_prints_contract_line():
    line = json.loads(_run().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "extrapolated" in cb["sample"] and cb["value"] == line["value"]
    assert line["config"]["workload"].startswith("2000 synthetic Gaussians") and line["scaling"] == "strong"


def test_reference_arm_other_ranks_exit_quietly():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
