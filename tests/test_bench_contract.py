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
    # default = the configuration BASELINE.json's metric is quoted on (direction-optimised BFS, RMAT-26)
    assert j["metric"] == "MTEPS (bfs_do_rmat26)" and j["steps"] == 2 and j["warmup"] == 1 and j["n_gpus"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] == j["e2e"]["value"]
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["config"]["workload"].startswith("BFS direction-optimised, RMAT-26")
    assert j["config"]["numerator"].startswith("sum of out-degrees") and j["scaling"] == "weak"
    # N > 1 names the partitioned traversal of the same graph (strong scaling); the CPU arm times the same BFS
    j8 = json.loads(run_bench("--gpus", "8")[0])
    assert j8["metric"] == "MTEPS (bfs_part_rmat26)" and j8["scaling"] == "strong" and j8["n_gpus"] == 8
    assert j8["config"]["edges_touched_per_step"] == j["config"]["edges_touched_per_step"]
    j = json.loads(run_bench("--workload", "bfs_push_rmat22")[0])
    assert j["metric"] == "MTEPS (bfs_push_rmat22)" and j["config"]["workload"].startswith("BFS push")
    # other algorithms of the path
    j = json.loads(run_bench("--workload", "sssp_rmat24")[0])
    assert j["metric"] == "MTEPS (sssp_rmat24)" and j["dtype"] == "f32" and j["value"] > 0


def test_reference_arm_other_ranks_are_silent(built):
    assert run_bench(env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []


def test_reference_gpu_leg_plumbing(tmp_path, monkeypatch):
    """bench.py --reference-gpu (the unmodified reference GPU kernels as a same-GPU baseline) cannot run without
    a GPU, but its plumbing can: the graph file must be in the reference's .csr layout (formats/csr.hxx:193-228)
    and the binary's JSON line must come back as MTEPS.  A stand-in executable checks the file it is given."""
    import importlib.util
    import numpy as np
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fake_root = tmp_path / "repo"
    (fake_root / "oracle" / "_ref").mkdir(parents=True)
    exe = fake_root / "oracle" / "_ref" / "gunrock_ref_gpu"
    exe.write_text(f"""#!{sys.executable}
import json, sys
import numpy as np
alg, path, src, runs, lb = sys.argv[1:6]
raw = np.fromfile(path, np.int32)
n, m, nnz = raw[:3]
ro = raw[3:3 + n + 1]
ci = raw[4 + n:4 + n + nnz]
vals = raw[4 + n + nnz:].view(np.float32)
assert n == m and ro[0] == 0 and ro[-1] == nnz and len(ci) == nnz and len(vals) == nnz and ci.max() < n
print("some library chatter")
print(json.dumps({{"impl": "reference_gpu", "algorithm": alg, "load_balance": lb, "vertices": int(n), "edges": int(nnz),
                  "source": int(src), "errors": 0 if "validate" in sys.argv else -1, "ms": [9.0] + [2.0] * (int(runs) - 1)}}))
""")
    exe.chmod(0o755)
    monkeypatch.setattr(bench, "ROOT", str(fake_root))

    class FakeGraph:
        def download(self):
            ro = np.array([0, 2, 3, 3, 5], np.int32)
            return ro, np.array([1, 2, 0, 0, 1], np.int32), None

    out = bench.reference_gpu_leg(FakeGraph(), {"alg": "bfs"}, 0, edges_per_run=4000, runs=4)
    assert out["block_mapped"]["errors_vs_reference_cpu"] == 0 and out["merge_path"]["errors_vs_reference_cpu"] == -1
    assert out["block_mapped"]["ms_median"] == 2.0 and out["block_mapped"]["runs"] == 4      # first run dropped
    assert abs(out["block_mapped"]["mteps"] - 4000 / 2.0 / 1e3) < 1e-12
    monkeypatch.setattr(bench, "ROOT", str(tmp_path / "nowhere"))
    assert "unavailable" in bench.reference_gpu_leg(FakeGraph(), {"alg": "bfs"}, 0, 1)


def test_clock_sampler_child_process_protocol(monkeypatch):
    """N > 1 runs sample clocks from a CHILD process (bench.ClockSampler.start_child / stop_child): time-stamped
    lines, only those inside the timed region count, throttle bits become reason names, a region shorter than the
    polling period falls back to all samples, and a child that cannot start NVML yields 'no sampler' -- never an
    exception.  The child here is a stand-in that prints the same line format without NVML."""
    import time
    import bench

    fake = ("import sys,time\n"
            "print('max',1965,flush=True)\n"
            "i=0\n"
            "while True:\n"
            "    print(time.time(),1965-15*(i%2),0x4 if i%5==0 else 0,flush=True)\n"
            "    i+=1\n"
            "    time.sleep(float(sys.argv[2]))\n")
    monkeypatch.setattr(bench.ClockSampler, "CHILD", fake)
    s = bench.ClockSampler(0, period_s=0.005)
    s.start_child()
    assert s.kind == "nvml-child"
    time.sleep(0.3)                       # "warm-up": samples before the region must not count
    t0 = time.time()
    time.sleep(0.2)
    t1 = time.time()
    c = s.stop_child(t0, t1)
    assert c["sm_max_mhz"] == 1965.0 and c["sm_mhz"] in (1950.0, 1957.5, 1965.0)
    assert 5 <= c["samples"] < c["samples_total"] and c["reasons"] == ["sw_power_cap"]
    assert "inside the timed region" in c["source"]
    # a region no sample falls into: all samples of the run, and the note says so
    s = bench.ClockSampler(0, period_s=0.005)
    s.start_child()
    time.sleep(0.2)
    c = s.stop_child(0.0, 1.0)
    assert c["samples"] == c["samples_total"] > 0 and "no sample fell inside" in c["source"]
    # a child that dies at once (no NVML on this box: the real CHILD)
    monkeypatch.undo()
    s = bench.ClockSampler(0, period_s=0.005)
    s.start_child()
    c = s.stop_child(time.time(), time.time() + 1)
    assert c["sm_mhz"] is None or isinstance(c["sm_mhz"], float)      # None here; a float on a GPU box
    assert isinstance(c["reasons"], list)
