"""The bench's reference arm (`bench.py --impl reference`) on the CPU, with a small sample: one JSON line on stdout
and nothing else, the keys the driver reads, and ranks other than 0 leaving without work.  (The GPU arm needs a
device; `tools/final_pass.sh` runs it on the box.)"""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-batches", "1", "--cpu-rows", "16",
         "--no-kriging", "--no-kriging-1deg", "--no-stochastic", "--no-ellipse-fit"]


def run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *SMALL], capture_output=True, text=True,
                          env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    p = run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.strip().splitlines()
    assert len(lines) == 1, p.stdout[:500]                 # library / reference prints go to stderr
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "entries/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "1deg" in d["config"]["workload"]
    nb = d["notebook"]                                     # configs[0], whole, on the host cores
    assert "error" not in nb["cpu_baseline"], nb
    assert nb["unit"] == "fields/s" and nb["value"] > 0
    assert "from the committed golden" in nb["cpu_baseline"]["sample"]


def test_reference_arm_other_ranks_leave_without_work():
    p = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and p.stdout.strip() == ""
