"""bench.py's reference arm runs on the CPU (oracle/_ref or the oracle port), so its JSON contract can be
checked here: one line, the keys the driver reads, rank != 0 stays silent."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def run_bench(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "12",
                        "--steps", "2", "--warmup", "1", *extra], capture_output=True, text=True, timeout=600, env=e)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.strip()]


def test_reference_arm_line(built):
    lines = run_bench()
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "MTEPS" and j["higher_is_better"] is True
    assert j["metric"] == "MTEPS (bfs_push_rmat22)" and j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] == j["e2e"]["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["config"]["workload"].startswith("BFS push")
    # other algorithms of the path
    j = json.loads(run_bench("--workload", "sssp_rmat24")[0])
    assert j["metric"] == "MTEPS (sssp_rmat24)" and j["dtype"] == "f32" and j["value"] > 0


def test_reference_arm_other_ranks_are_silent(built):
    assert run_bench(env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
