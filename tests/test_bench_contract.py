"""CPU check of the bench.py contract that does not need a GPU: the reference arm (`--impl reference`) prints exactly
ONE JSON line on stdout -- whatever libraries write to fd 1 is diverted to stderr -- with the keys the driver reads;
under a multi-rank launch only rank 0 prints.  (The B200 arm needs a device; its line has the same skeleton plus
`roofline`, `clocks`, `gpu_launches`, checked on the GPU box by the driver itself.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-n", "20000", "--nlist", "64"]


def run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True,
                          env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["metric"] == "IVF-Flat QPS @10M x 768D nprobe=32 k=10" and d["value"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("IVF-Flat 10M x 768D L2 nlist=64")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_is_silent_on_ranks_other_than_zero():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])
