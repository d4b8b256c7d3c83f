"""bench.py's reference arm prints ONE JSON line with the contract's keys (runs the oracle on the
host CPU: no GPU needed on the tiny configuration).  The GPU arm's line is checked by the driver."""
import json
import os
import subprocess
import sys

from util import ROOT


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny",
                        "--workload", "C1-tiny", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train samples/s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 2 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("C1-tiny")
    # the CPU arm's config says what ONE of its steps covers (it used to copy the GPU arm's)
    assert d["config"]["step_sample"] == {"per_step_batch": 1, "accum": 1, "of": {"per_gpu_batch": 2, "accum": 2}}
    assert "CPU arm: each step = 1 sample x accum 1" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_on_other_ranks_exits_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--model", "tiny",
                        "--workload", "C1-tiny", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]
