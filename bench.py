#!/usr/bin/env python
"""bench.py -- MTEPS of the frontier hot path on synthetic RMAT graphs (BASELINE.json metric).

    python bench.py [--gpus 1] --steps K --warmup W           # our CUDA path: ONE JSON line
    torchrun ... bench.py --gpus N --steps K --warmup W       # N > 1: the same traversal, 1-D partitioned
    python bench.py --impl reference ...                      # the reference's CPU path, same config

The headline workload is the one BASELINE.json's metric is quoted on: direction-optimised BFS on RMAT-26
(scale 26, 16 directed edges per vertex) from the highest-degree vertex.
  N = 1 : the fused single-GPU enactor (`bfs_do_rmat26`).
  N > 1 : ONE traversal over the graph 1-D (cyclic) vertex-cut across the N ranks (`bfs_part_rmat26`), the
          per-level frontier exchange done by our kernels over NVLink peer memory (default) or by NCCL
          (`--exchange nccl`; its time is also reported beside the default) -> "scaling": "strong".
A "step" is one full traversal over the graph already resident in HBM.
  value    : MTEPS with the SAME numerator at every N and for the CPU arm: the sum of out-degrees of the
             reached vertices (= the edges the reference's CPU BFS traverses, graph500 convention) per second;
             `config.edges_inspected_per_step` is what the direction-optimised run actually read.
  e2e      : same metric through the C-ABI call with HOST buffers (source in, the V x 4 B result copied to
             pinned host memory inside the timed region)
  roofline : the dominant kernel class's algorithmic bytes (4 B per column index actually read; 8 B for SSSP)
             over its device time (per-level CUDA events recorded on the launching stream by the enactor)
  cpu_baseline : the reference's own CPU validator (oracle/_ref, compiled from /root/reference; the oracle's
             C port where that binary is absent) on this box's host cores, with a full-size parity check
  configs  : (N = 1, default run) the other BASELINE.json configurations as sub-records of the same line --
             bfs_push_rmat22 (merge_path), sssp_rmat24 (block_mapped), pr_lj -- each with value / roofline /
             cpu_baseline / full-size parity.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "bfs_push_rmat22": dict(alg="bfs", scale=22, ef=16, seed=0x5EED22, mirror=True, fold=0, weights=0,
                            lb="merge_path", direction="forward",
                            desc="BFS push, RMAT-22 ef16 symmetrised+dedup, merge_path advance + in-kernel compact filter"),
    "bfs_do_rmat22": dict(alg="bfs", scale=22, ef=16, seed=0x5EED22, mirror=True, fold=0, weights=0,
                          lb="block_mapped", direction="optimized",
                          desc="BFS direction-optimised, RMAT-22 ef16"),
    "sssp_rmat24": dict(alg="sssp", scale=24, ef=16, seed=0x5EED24, mirror=True, fold=0, weights=1,
                        lb="block_mapped", direction="forward",
                        desc="SSSP fp32 weights 1..63, RMAT-24 ef16, block_mapped advance"),
    "pr_lj": dict(alg="pr", scale=23, ef=0, pairs=82_000_000, seed=0x5EED4C, mirror=False, fold=4_847_571,
                  weights=0, lb="block_mapped", direction="backward",
                  desc="PageRank pull, soc-LiveJournal1-shaped RMAT (4.85M V, ~69M E directed)"),
    "bfs_do_rmat26": dict(alg="bfs", scale=26, ef=8, seed=0x5EED26, mirror=True, fold=0, weights=0,
                          lb="block_mapped", direction="optimized",
                          desc="BFS direction-optimised, RMAT-26 with 16 directed edges/vertex (int32-safe)"),
    # BASELINE.json configs[4]: 1-D vertex partition over the ranks + per-level frontier exchange (strong scaling)
    "bfs_part_rmat26": dict(alg="bfs", scale=26, ef=8, seed=0x5EED26, mirror=True, fold=0, weights=0,
                            lb="block_mapped", direction="optimized", partitioned=True,
                            desc="BFS direction-optimised, RMAT-26 (16 directed edges/vertex), 1-D cyclic vertex "
                                 "partition across the ranks, per-level frontier exchange"),
    "bfs_part_rmat22": dict(alg="bfs", scale=22, ef=16, seed=0x5EED22, mirror=True, fold=0, weights=0,
                            lb="block_mapped", direction="optimized", partitioned=True,
                            desc="BFS direction-optimised, RMAT-22 ef16, 1-D partition + per-level frontier exchange"),
}
HEADLINE_1, HEADLINE_N = "bfs_do_rmat26", "bfs_part_rmat26"
SUB_CONFIGS = ["bfs_push_rmat22", "sssp_rmat24", "pr_lj"]
INT_MAX = 2**31 - 1
FLT_MAX = float(np.finfo(np.float32).max)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    In-process NVML (pynvml) from a host thread every 10 ms: two cheap queries per sample.  The first version spawned
    `nvidia-smi --query-gpu=<9 fields> -lms 100`; every one of its queries takes the driver's lock for milliseconds,
    and one in ten timed runs of a few milliseconds came out twice as long (profiles/r2_bench_matrix.md).  nvidia-smi
    remains the fallback when NVML cannot be loaded."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index, period_s=0.01):
        self.idx = device_index
        self.period = period_s
        self.lines = []
        self.proc = None
        self.samples = []          # (sm_mhz, reasons bit mask)
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.thread = None
        self.kind = None

    def start(self):
        if os.environ.get("B2G_BENCH_NO_SAMPLER"):
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.idx
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:   # NVML counts physical devices
                try:
                    idx = int(vis.split(",")[self.idx])
                except (ValueError, IndexError):
                    pass
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop_flag.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(get_reasons(h))))
                    except Exception:
                        pass
                    self.stop_flag.wait(self.period)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.kind = "nvml"
            return
        except Exception:
            self.kind = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            self.kind = "nvidia-smi"
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    # ---- NVML from a SEPARATE process (multi-GPU runs): an NVML query made inside a rank's own process stalls that
    # rank's level loop for milliseconds; a child process polls the same two values and prints time-stamped lines,
    # the parent keeps those that fall inside the timed region.
    CHILD = ("import sys,time,pynvml\n"
             "pynvml.nvmlInit()\n"
             "h=pynvml.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))\n"
             "g=getattr(pynvml,'nvmlDeviceGetCurrentClocksEventReasons',None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons\n"
             "print('max',pynvml.nvmlDeviceGetMaxClockInfo(h,pynvml.NVML_CLOCK_SM),flush=True)\n"
             "while True:\n"
             "    print(time.time(),pynvml.nvmlDeviceGetClockInfo(h,pynvml.NVML_CLOCK_SM),int(g(h)),flush=True)\n"
             "    time.sleep(float(sys.argv[2]))\n")

    def start_child(self):
        """Spawn the polling child now (its start-up takes ~0.2 s: call this before the warm-up steps)."""
        idx = self.idx
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                idx = int(vis.split(",")[self.idx])
            except (ValueError, IndexError):
                pass
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", self.CHILD, str(idx), str(self.period)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            self.kind = "nvml-child"
            # Wait until the child is SAMPLING (its first lines arrive) before any timed work starts: nvmlInit
            # attaches every GPU of the box and holds the driver's locks for up to seconds -- call X (2 ranks) had
            # the child still initialising during the 17 ms timed region: no sample at all and one traversal of
            # 8.6 ms among 0.75 ms ones.  Bounded: a child that cannot load NVML exits and the wait ends.
            deadline = time.time() + 20.0
            while time.time() < deadline and len(self.lines) < 3 and self.proc.poll() is None:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def stop_child(self, t_begin, t_end):
        """Samples whose time stamps fall inside [t_begin, t_end] (host clock); ends the child."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sampler"]}
        time.sleep(2.5 * self.period)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        mx, sm, reasons, total, inside, every = None, [], set(), 0, [], []
        for ln in self.lines:
            f = ln.split()
            try:
                if f[0] == "max":
                    mx = float(f[1])
                    continue
                t, mhz, mask = float(f[0]), float(f[1]), int(f[2])
            except (ValueError, IndexError):
                continue
            total += 1
            every.append((mhz, mask))
            if t_begin <= t <= t_end:
                inside.append((mhz, mask))
        note = "samples inside the timed region"
        if not inside and every:    # a region shorter than the polling period: fall back to the whole run's samples
            inside, note = every, "no sample fell inside the timed region: all samples of the run (warm-up included)"
        for mhz, mask in inside:
            sm.append(mhz)
            for bit, nm in names.items():
                if mask & bit:
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "samples_total": total, "reasons": sorted(reasons) if sm else ["no sampler"],
                "source": "NVML polled every %g ms by a child process; %s" % (self.period * 1e3, note)}

    def stop(self):
        if self.kind == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            reasons = set()
            for _, mask in self.samples:
                for bit, nm in names.items():
                    if mask & bit:
                        reasons.add(nm)
            sm = [x for x, _ in self.samples]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                    "samples": len(sm), "reasons": sorted(reasons), "source": "NVML, in-process, every 10 ms"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sampler"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------
# host side: the checker's graph build and the CPU runs (oracle / oracle/_ref -- never the product)
# ---------------------------------------------------------------------------------------------
def host_graph(wl):
    """The workload's CSR built on the HOST by the checker's generator with all host threads (reference arm).
    Bit-identical to the device generator (tests/test_gpu_parity) and to the serial builder (tests/test_oracle)."""
    import oracle
    n_pairs = wl.get("pairs") or wl["ef"] * (1 << wl["scale"])
    ro, ci = oracle.rmat_csr_parallel(wl["scale"], n_pairs, wl["seed"], wl["mirror"], wl["fold"])
    w = oracle.edge_weights_parallel(wl["seed"] + 1, ro, ci, wl["weights"] == 2) if wl["weights"] else None
    return ro, ci, w


def cpu_run_factory(wl, ro, ci, w):
    """Returns (run(source) -> (result, ms), kind, cores).  Prefers the compiled reference."""
    import oracle
    alg = wl["alg"]
    if alg in ("bfs", "sssp") and oracle.ref_available():
        g = oracle.RefGraph(ro, ci, w)
        fn = g.bfs if alg == "bfs" else g.sssp
        return (lambda s: fn(s)), "reference", 1
    if alg == "bfs":
        def run(s):
            t = time.perf_counter()
            r = oracle.bfs(ro, ci, s)
            return r, (time.perf_counter() - t) * 1e3
    elif alg == "sssp":
        def run(s):
            t = time.perf_counter()
            r = oracle.sssp(ro, ci, w, s)
            return r, (time.perf_counter() - t) * 1e3
    else:
        def run(s):
            t = time.perf_counter()
            r = oracle.pr(ro, ci, w, 0.85, 1e-6)
            return r, (time.perf_counter() - t) * 1e3
    return run, "port", 1


def reached_degree_sum(wl, ro, result, iters=None):
    """The TEPS numerator shared by every arm: sum of out-degrees of the reached vertices (BFS, SSSP);
    E x iterations for PageRank (performance.hxx:225-229 counts edges visited)."""
    deg = np.diff(ro).astype(np.int64)
    if wl["alg"] == "bfs":
        return int(deg[np.asarray(result) < INT_MAX].sum())
    if wl["alg"] == "sssp":
        return int(deg[np.asarray(result) < FLT_MAX].sum())
    return int(deg.sum()) * int(iters)


def bench_source(ro):
    return int(np.diff(ro).argmax())


def parity(wl, got, want):
    """Full-size check of the product's result against the checker's (bit-exact for BFS / SSSP, 1e-6 relative
    for PageRank).  Returns (ok, detail)."""
    if wl["alg"] == "pr":
        want_p = np.asarray(want[0], np.float64)
        rel = np.abs(np.asarray(got, np.float64) - want_p) / np.maximum(np.abs(want_p), 1e-300)
        return bool(rel.max() <= 1e-6), {"max_rel_err": float(rel.max()), "tolerance": 1e-6}
    a = np.asarray(got).view(np.uint32)
    b = np.asarray(want).view(np.uint32)
    bad = int((a != b).sum())
    return bad == 0, {"mismatches": bad, "rule": "bit-exact"}


def reference_gpu_leg(G, wl, src, edges_per_run, runs=5):
    """Times oracle/_ref/gunrock_ref_gpu (the unmodified reference GPU path) on the bench graph."""
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "gunrock_ref_gpu")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/gunrock_ref_gpu not built (needs /root/reference at build time)"}
    ro, ci, w = G.download()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "graph.csr")
        with open(path, "wb") as f:               # formats/csr.hxx:193-228: rows, cols, nnz, offsets, indices, values
            np.array([len(ro) - 1, len(ro) - 1, len(ci)], np.int32).tofile(f)
            ro.tofile(f)
            ci.tofile(f)
            (w if w is not None else np.ones(len(ci), np.float32)).tofile(f)
        out = {}
        for lb in ("block_mapped", "merge_path"):
            cmd = [exe, wl["alg"], path, str(src), str(runs), lb]
            if wl["alg"] != "pr" and len(ci) <= 200_000_000 and lb == "block_mapped":
                cmd.append("validate")            # its own CPU validator, once
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
            except subprocess.TimeoutExpired:
                out[lb] = {"error": "timed out after 900 s"}
                continue
            if r.returncode != 0:
                out[lb] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            j = json.loads(r.stdout.strip().splitlines()[-1])
            ms = sorted(j["ms"][1:] or j["ms"])   # first run pays the lazy module load
            out[lb] = {"ms_median": statistics.median(ms), "ms_best": ms[0], "runs": len(j["ms"]),
                       "errors_vs_reference_cpu": j["errors"],
                       "mteps": (edges_per_run / statistics.median(ms) / 1e3) if edges_per_run else None}
    return {"kind": "unmodified reference GPU kernels, nvcc sm_100a, -include oracle/ref_gpu_fix.h, SM_TARGET=90",
            "timed_region": "the reference enactor's own timer (enactor.hxx:266-288), reset excluded",
            "numerator": "sum of out-degrees of the reached vertices (bfs / sssp); E x our iteration count (pr)",
            **out}


def run_reference(args, wl, name):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    t0 = time.time()
    ro, ci, w = host_graph(wl)
    build_s = time.time() - t0
    run, kind, cores = cpu_run_factory(wl, ro, ci, w)
    src = bench_source(ro)
    # A step of this arm is one FULL run of the reference's CPU implementation (11-17 s for BFS on RMAT-26): the
    # sample is bounded by running as many of the requested W + K runs as fit a wall-clock budget (at least one
    # timed run); MTEPS is a rate, so fewer runs change its noise, not its meaning.
    budget_s = float(os.environ.get("B2G_REFERENCE_BUDGET_S", "100"))
    t_first = time.time()
    res, t = run(src)
    per_run_s = max(time.time() - t_first, 1e-3)
    ms = []
    n_warm = min(args.warmup, max(1, int(0.2 * budget_s / per_run_s))) if args.warmup > 0 else 0
    if n_warm == 0:
        ms.append(t)                      # no warm-up requested: the first run is a timed one
    for _ in range(max(n_warm - 1, 0)):
        run(src)
    n_timed = min(args.steps, max(1, int((budget_s - n_warm * per_run_s) / per_run_s)))
    while len(ms) < n_timed:
        res, t = run(src)
        ms.append(t)
    et = reached_degree_sum(wl, ro, res[0] if wl["alg"] == "pr" else res, res[1] if wl["alg"] == "pr" else None)
    total_ms = sum(ms)
    value = et * len(ms) / total_ms / 1e3
    line = {"impl": "reference", "metric": f"MTEPS ({name})", "value": value, "unit": "MTEPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "steps_timed": len(ms), "warmup_run": n_warm,
            "ms_per_step": total_ms / max(len(ms), 1), "higher_is_better": True,
            "scaling": "strong" if wl.get("partitioned") else "weak",
            "vs_baseline": None, "dtype": "int32" if wl["alg"] == "bfs" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "vertices": int(len(ro) - 1), "edges": int(len(ci)),
                       "source": src, "edges_touched_per_step": et,
                       "numerator": "sum of out-degrees of the reached vertices"},
            "cpu_baseline": {"value": value, "unit": "MTEPS", "cores": cores, "kind": kind,
                             "sample": f"{len(ms)} full {wl['alg']} run(s) (of {args.steps} requested; bounded by a "
                                       f"{budget_s:.0f} s budget, B2G_REFERENCE_BUDGET_S) from the bench source, "
                                       f"validator's own timer; host cores available: {os.cpu_count()}; "
                                       f"the graph was built on the host with {oracle.num_threads()} threads "
                                       f"in {build_s:.1f} s (untimed)"},
            "e2e": {"value": value, "unit": "MTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "setup_s": round(time.time() - t0 - total_ms / 1e3, 1)}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# one workload on one GPU through the C ABI
# ---------------------------------------------------------------------------------------------
def load_traffic(name, wl):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this very configuration
    (profiles/r2_traffic.json, falling back to round 1's file) -- never measured under the profiler here."""
    for fn in ("r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", fn)) as f:
                entry = json.load(f).get(f"{name}/{wl['lb']}")
        except (OSError, ValueError):
            continue
        if entry:
            per = [l["dram_read_bytes"] + l["dram_write_bytes"] for l in entry["launches"]]
            return sum(per) / len(per), {"source": entry["source"], "launches": entry["launches"],
                                         "note": "mean over the capture's launches; compare with their algorithmic_bytes"}
    return None, None


def bench_single(args, name, wl, steps, warmup, local, cpu_baseline=True, background_cpu=False, extras=True):
    """Returns (record, finish) -- `record` is the JSON object of this workload; when `background_cpu` is set
    the CPU baseline + parity run on a host thread and `finish()` joins it and fills them in."""
    import torch
    import gunrock_b200 as gb

    t_setup = time.time()
    n_pairs = wl.get("pairs") or wl["ef"] * (1 << wl["scale"])
    G = gb.graph_t.rmat(wl["scale"], n_pairs, wl["seed"], mirror=wl["mirror"], fold_vertices=wl["fold"],
                        weights=wl["weights"], weight_seed=wl["seed"] + 1)
    src, src_deg = G.max_degree_vertex()
    if wl["alg"] == "pr" or wl["direction"] != "forward":
        G.build_transpose()
    stream = torch.cuda.Stream()
    opt = gb.options_t(advance_load_balance=getattr(gb.load_balance_t, wl["lb"]),
                       advance_direction=getattr(gb.advance_direction_t, wl["direction"]),
                       filter_algorithm=gb.filter_algorithm_t.compact, enable_filter=True,
                       hub_threshold=args.hub_threshold, ctas_per_sm=args.ctas_per_sm,
                       stream=stream.cuda_stream)
    V = G.n_vertices
    out_dtype = torch.int32 if wl["alg"] == "bfs" else torch.float32
    d_out = torch.empty(V, dtype=out_dtype, device="cuda")
    h_out = torch.empty(V, dtype=out_dtype).pin_memory()
    bytes_per_edge = 8 if wl["alg"] == "sssp" else 4

    def step(out):
        if wl["alg"] == "bfs":
            return gb.bfs(G, src, out, options=opt)
        if wl["alg"] == "sssp":
            return gb.sssp(G, src, out, options=opt)
        return gb.pr(G, out, 0.85, 1e-6, options=opt)

    def timed(out, k):
        """K steps bracketed by synchronize; CUDA events on the launching stream."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        agg = dict(inspected=0, launches=0, run_ms=[], lv=[])
        torch.cuda.synchronize()
        t_wall = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(k):
                st = step(out)
                agg["inspected"] += st.edges_touched
                agg["launches"] += st.kernel_launches
                agg["run_ms"].append(float(st.elapsed_ms))   # the library's own events around enact()
                agg["lv"].append((list(st.level_direction), list(st.level_edges), list(st.level_kernel_ms)))
            e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), (time.perf_counter() - t_wall) * 1e3, agg, st

    for _ in range(warmup):
        step(d_out)
    ms, wall_ms, agg, last = timed(d_out, steps)
    step(h_out)
    ms_e2e, _, agg_e2e, _ = timed(h_out, steps)
    result = h_out.numpy().copy()

    # ---- numerator (same for every arm): needs the row offsets on the host ------------------------
    ro, ci, w = G.download()
    iters = last.iterations
    touched = reached_degree_sum(wl, ro, result, iters)
    value = touched * steps / ms / 1e3                    # edges / ms / 1000 = MTEPS (performance.hxx:225-229)
    e2e = touched * steps / ms_e2e / 1e3
    peak, peak_kind = peaks()

    # ---- roofline of the dominant kernel class ----------------------------------------------------
    # per-level CUDA events recorded by the enactor on the launching stream; a level's kernel time covers its
    # advance / sweep launches.  BFS direction-optimised: the pull sweeps dominate; otherwise the advances.
    dom_dir = 1 if (wl["alg"] == "bfs" and wl["direction"] != "forward") else 0
    kb = km = 0.0
    kl = 0
    other_b = other_ms = 0.0
    for dirs, edges, kms in agg["lv"]:
        for i, (e, t) in enumerate(zip(edges, kms)):
            d = dirs[i] if i < len(dirs) else 0
            if wl["alg"] != "bfs" or d == dom_dir:
                kb += e * bytes_per_edge
                km += t
                kl += 1
            else:
                other_b += e * bytes_per_edge
                other_ms += t
    ach = kb / (km * 1e-3) / 1e9 if km > 0 else 0.0
    lbk = "merge_path" if wl["lb"] == "merge_path" else "binned"
    kernel = {"bfs": "bfs_pull_first_kernel + bfs_pull_rest_kernel (pull levels)" if dom_dir else
              "advance_%s_kernel<bfs_claim_op>" % lbk,
              "sssp": "advance_%s_kernel<sssp_relax_op>%s" % (lbk, " + advance_hub_kernel" if lbk == "binned" else ""),
              "pr": "pr_pull_tile_kernel"}[wl["alg"]]
    traffic, traffic_detail = load_traffic(name, wl) if not args.scale else (None, None)
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_detail": traffic_detail, "peak_kind": peak_kind, "kernel": kernel,
                "bytes_per_edge": bytes_per_edge, "launches": kl, "kernel_ms_total": km,
                "algorithmic_bytes_total": kb,
                "whole_step": {"achieved": agg["inspected"] * bytes_per_edge / (ms * 1e-3) / 1e9,
                               "frac": agg["inspected"] * bytes_per_edge / (ms * 1e-3) / 1e9 / peak}}
    if other_ms > 0:
        roofline["other_levels"] = {"kernel": "advance kernels of the push levels", "achieved":
                                    other_b / (other_ms * 1e-3) / 1e9, "kernel_ms_total": other_ms}
    # the metric's own ceiling is peak / (bytes per touched edge); a direction-optimised run
    # touches (credits) every reached vertex's edges but inspects few of them, so its value may pass this ceiling
    ceiling = peak * 1e9 / bytes_per_edge / 1e6
    roofline["metric_ceiling"] = {"mteps": ceiling, "value_frac": value / ceiling,
                                  "rule": "measured HBM peak / %d B per touched edge (SURVEY.md 8d)" % bytes_per_edge}
    if dom_dir and len(agg["lv"]):
        # secondary full-traffic model of the pull levels (SURVEY.md 8d "report, don't grade"): what K1 / K2 must
        # move however few edges they inspect -- 8 B of stored in-neighbours per unvisited vertex that has
        # in-edges, 4 B per label written, the four bitmaps (visited r/w, next w, retry w), 4 B per inspected
        # edge beyond the two stored ones
        dirs, edges, kms = agg["lv"][-1]
        fr = list(last.level_frontier)[:len(edges)]
        deg = np.diff(ro)
        with_in = int(np.count_nonzero(deg)) if wl.get("symmetric", True) else len(deg)
        seen, model_b, model_ms = 0, 0.0, 0.0
        for i, (e, t) in enumerate(zip(edges, kms)):
            seen += fr[i]
            if (dirs[i] if i < len(dirs) else 0) == 1:
                nxt = fr[i + 1] if i + 1 < len(fr) else 0
                unvisited = max(with_in - seen, 0)
                model_b += 8.0 * unvisited + 4.0 * nxt + 4.0 * (len(deg) / 8.0) + 4.0 * max(e - 2 * unvisited, 0)
                model_ms += t
        if model_ms > 0:
            roofline["pull_levels_full_traffic_model"] = {
                "bytes_per_run": model_b, "kernel_ms_per_run": model_ms,
                "achieved": model_b / (model_ms * 1e-3) / 1e9, "frac": model_b / (model_ms * 1e-3) / 1e9 / peak,
                "rule": "8 B x unvisited vertices with in-edges + 4 B x labels written + 4 bitmaps of V/8 B + "
                        "4 B x edges inspected beyond the two stored per vertex; last timed run"}

    r_sorted = sorted(agg["run_ms"])
    runs = {"best_ms": r_sorted[0], "median_ms": statistics.median(r_sorted), "worst_ms": r_sorted[-1],
            "best_mteps": touched / r_sorted[0] / 1e3, "median_mteps": touched / statistics.median(r_sorted) / 1e3,
            "region": "CUDA events around the enactor loop of each run (enactor.hxx:266-288)"}
    sum_kernel_ms = sum(sum(k) for _, _, k in agg["lv"]) / steps
    record = {
        "metric": f"MTEPS ({name})", "value": value, "unit": "MTEPS", "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32" if wl["alg"] == "bfs" else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "vertices": V, "edges": G.n_edges, "source": src,
                   "source_degree": src_deg, "load_balance": wl["lb"], "direction": wl["direction"],
                   "filter": "compact (fused into advance)",
                   "numerator": "sum of out-degrees of the reached vertices (PageRank: E x iterations)",
                   "edges_touched_per_step": touched, "edges_inspected_per_step": agg["inspected"] // steps,
                   "inspected_mteps": agg["inspected"] / ms / 1e3,
                   "l2_policy": "inputs larger than L2 (column indices %.0f MB > 126 MB)" % (G.n_edges * 4 / 1e6),
                   "levels": last.iterations, "level_direction": last.level_direction,
                   "level_frontier": last.level_frontier, "level_edges": last.level_edges[:last.iterations],
                   "level_kernel_ms": [round(x, 4) for x in last.level_kernel_ms[:last.iterations]],
                   "kernel_ms_per_step": sum_kernel_ms, "outside_kernels_frac": 1.0 - sum_kernel_ms / (ms / steps),
                   "runs": runs,
                   "experimental": {k: os.environ[k] for k in os.environ if k.startswith("B2G_")}},
        "e2e": {"value": e2e, "unit": "MTEPS", "h2d_bytes_per_step": 4, "d2h_bytes_per_step": V * 4,
                "ms_per_step": ms_e2e / steps},
        "gpu_launches": agg["launches"], "roofline": roofline, "cpu_baseline": None, "wall_ms": wall_ms,
    }

    # 16 random sources of degree > 0, device-resident results; never takes the line down
    if extras and wl["alg"] in ("bfs", "sssp"):
        try:
            picks = np.random.default_rng(1).choice(np.flatnonzero(np.diff(ro) > 0), 16, replace=False)
            per = []
            fn = gb.bfs if wl["alg"] == "bfs" else gb.sssp
            with torch.cuda.stream(stream):
                for s16 in picks:
                    st16 = fn(G, int(s16), d_out, options=opt)
                    per.append(st16.elapsed_ms)
            record["config"]["runs"]["random16_mean_ms"] = float(np.mean(per))
            record["config"]["runs"]["random16_max_ms"] = float(np.max(per))
        except Exception as ex:
            record["config"]["runs"]["random16_error"] = str(ex)
        step(d_out)  # leave the bench source's result in place

    if args.reference_gpu:
        try:
            record["reference_gpu"] = reference_gpu_leg(G, wl, src, touched)
            for lb in ("block_mapped", "merge_path"):
                r = record["reference_gpu"].get(lb)
                if r and r.get("ms_median"):
                    r["ours_over_reference_gpu"] = r["ms_median"] / runs["median_ms"]
        except Exception as ex:
            record["reference_gpu"] = {"error": str(ex)}
    G.close()
    del G
    record["setup_s"] = round(time.time() - t_setup, 1)

    # ---- CPU baseline on this box's host cores + full-size parity -----------------------------------
    def cpu_leg():
        try:
            run, kind, cores = cpu_run_factory(wl, ro, ci, w)
            res, cms = run(src)
            et = reached_degree_sum(wl, ro, res[0] if wl["alg"] == "pr" else res, res[1] if wl["alg"] == "pr" else None)
            cpu = {"value": et / cms / 1e3, "unit": "MTEPS", "cores": cores, "kind": kind,
                   "sample": f"1 full {wl['alg']} run from the bench source ({cms / 1e3:.1f} s, validator's own "
                             f"timer); host cores available: {os.cpu_count()}"}
            ok, detail = parity(wl, result, res)
            if wl["alg"] == "pr":
                detail["iterations_equal"] = bool(int(res[1]) == int(iters))
                ok = ok and detail["iterations_equal"]
            cpu["parity_full_size"] = ok
            cpu["parity_detail"] = detail
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "MTEPS", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
        record["cpu_baseline"] = cpu

    if not cpu_baseline:
        return record, (lambda: record)
    if background_cpu:
        th = threading.Thread(target=cpu_leg, daemon=True)   # ctypes releases the GIL during the CPU run
        th.start()

        def finish():
            th.join()
            return record
        return record, finish
    cpu_leg()
    return record, (lambda: record)


# ---------------------------------------------------------------------------------------------
# N > 1: ONE traversal over the 1-D partitioned graph (BASELINE.json configs[4], strong scaling)
# ---------------------------------------------------------------------------------------------
def run_partitioned(args, wl, name, rank, world, local):
    import torch
    import torch.distributed as dist
    import gunrock_b200 as gb
    from gunrock_b200 import multi_gpu as mg

    n_pairs = wl["ef"] * (1 << wl["scale"])
    G = mg.PartitionedGraph.rmat(wl["scale"], n_pairs, wl["seed"], world, rank, mirror=wl["mirror"])
    comm = mg.TorchDistComm()
    v, d = G.max_degree_vertex()
    key = torch.tensor([(d << 32) | (0x7fffffff - v)], dtype=torch.int64, device="cuda")
    dist.all_reduce(key, op=dist.ReduceOp.MAX)
    src = 0x7fffffff - int(key.item() & 0xffffffff)
    total_edges = comm.all_reduce_sum([G.n_local_edges], "cuda")[0]
    opt = gb.options_t(advance_load_balance=getattr(gb.load_balance_t, wl["lb"]),
                       hub_threshold=args.hub_threshold, ctas_per_sm=args.ctas_per_sm)
    eng = mg.CudaRankEngine(G, opt)
    direction = getattr(gb.advance_direction_t, wl["direction"])
    mg.p2p_connect(eng, comm)

    def step_p2p():   # the kernels exchange frontiers over NVLink peer memory (bfs_p2p.cuh)
        return mg.bfs_rank_p2p(eng, src, total_edges, direction)

    nccl_ready = [False]

    def ensure_nccl():  # the C++ level loop's own communicator (ncclCommInitRank inside the library)
        if not nccl_ready[0]:
            mg.nccl_connect(eng, comm)
            nccl_ready[0] = True

    def step_nccl():  # ncclSend/Recv all-to-all, ncclAllGather, ncclAllReduce -- enqueued by the C++ level loop
        ensure_nccl()
        return mg.bfs_rank_nccl(eng, src, total_edges, direction)

    primary, other = (step_p2p, step_nccl) if args.exchange == "p2p" else (step_nccl, step_p2p)

    def timed(step, k, copy_out=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        inspected = 0
        per_run.clear()
        for _ in range(k):
            dloc, st = step()
            inspected += st.edges_touched
            per_run.append(float(st.elapsed_ms))      # the library's own events around this rank's level loop
            if copy_out is not None:
                copy_out.copy_(dloc, non_blocking=False)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), inspected, dloc, st

    per_run = []

    def run_spread():
        r = sorted(per_run)
        return {"best_ms": r[0], "median_ms": statistics.median(r), "worst_ms": r[-1],
                "region": "rank 0, CUDA events around the level loop of each run"} if r else None

    warm = max(args.warmup, 3)
    # 10 ms, the period of the N = 1 sampler: 1-2 samples inside a 20-step region of ~20 ms (call Y, 5 ms: 3 inside,
    # the sampled region 1.13 ms / step against 1.01 un-sampled -- per-run medians 0.835 / 0.818); a region no sample
    # falls into reports all samples the child took (they start before the warm-up steps: the same load) and says so
    sampler = ClockSampler(local, period_s=0.010)
    if rank == 0 and not os.environ.get("B2G_BENCH_NO_SAMPLER"):
        sampler.start_child()      # returns once the child is sampling (its NVML start-up can take seconds)
    dist.barrier()                 # the other ranks must not enter a traversal (bounded device-side spins) before that
    for _ in range(warm):
        primary()
    # Clocks at N > 1.  An NVML query issued from a rank's OWN process stalls that rank's level loop for
    # milliseconds (call U, 2 ranks: per-run median 0.76 ms, but the one run in ten that met a 10 ms in-process
    # sample took 2.3 ms with the peer-memory exchange and 16.8 ms with NCCL; without any sampler best / median /
    # worst were 0.744 / 0.747 / 0.755 ms) -- the single-GPU enactor shows no such effect.  So the sampling during
    # the timed region is done by a CHILD process (started, and waited for until it is sampling, before the warm-up
    # steps: its nvmlInit stalls every rank's driver calls for the seconds it takes -- call X), and an identical
    # region is timed once more with no sampler at all and reported beside it (`config.unsampled`).
    t_begin = time.time()
    ms, inspected, dloc, st = timed(primary, args.steps)
    t_end = time.time()
    runs_primary = run_spread()
    clocks = sampler.stop_child(t_begin, t_end) if rank == 0 else None
    # e2e: the rank's slice of the result is copied to pinned host memory inside the timed region
    h = torch.empty(G.n_local, dtype=torch.int32).pin_memory()
    ms2, _, dloc, _ = timed(primary, args.steps, copy_out=h)
    ms_u, _, dloc, _ = timed(primary, args.steps)
    unsampled = {"ms_per_step": ms_u / args.steps, "runs": run_spread(),
                 "what": "the same %d-step region timed again with no clock sampler running" % args.steps}
    # the other exchange, same graph, same source, outside the headline's timed region
    other_rec = None
    try:
        for _ in range(3):
            other()
        k2 = min(args.steps, 10)
        ms_o, _, d_other, st_o = timed(other, k2)
        same = torch.tensor([int(torch.equal(dloc, d_other))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        other_rec = {"exchange": "nccl" if args.exchange == "p2p" else "p2p", "ms_per_step": ms_o / k2,
                     "steps": k2, "runs": run_spread(), "depths_equal_primary": bool(same.item())}
    except Exception as ex:
        other_rec = {"error": str(ex)[:300]}

    # ---- numerator + full-size parity over the GATHERED depths ----------------------------------------
    # vertex v lives on rank v % P at row v // P: gather the slices on rank 0 and interleave
    rows0 = mg.rows_of(G.n_global, world, 0)
    pad = torch.full((rows0,), INT_MAX, dtype=torch.int32, device="cuda")
    pad[:G.n_local] = dloc
    gathered = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, gathered, dst=0)
    line = None
    if rank == 0:
        full = torch.stack(gathered, dim=1).reshape(-1)[:G.n_global].contiguous()   # [row, rank] -> v = row*P + rank
        del gathered
        peak, peak_kind = peaks()
        par = {}
        touched = None
        try:
            # the same graph, whole, on this GPU: the single-GPU fused enactor's depths (bit-exact) ...
            G1 = gb.graph_t.rmat(wl["scale"], n_pairs, wl["seed"], mirror=wl["mirror"])
            d1 = torch.empty(G.n_global, dtype=torch.int32, device="cuda")
            gb.bfs(G1, src, d1, options=gb.options_t(advance_direction=gb.advance_direction_t.optimized))
            par["equal_single_gpu_enactor"] = bool(torch.equal(d1, full))
            ro, ci, w = G1.download()
            G1.close()
            del d1
            touched = reached_degree_sum(wl, ro, full.cpu().numpy())
            if not args.no_cpu_baseline:   # ... and the reference's CPU validator on this box's host cores
                run, kind, cores = cpu_run_factory(wl, ro, ci, None)
                res, cms = run(src)
                ok, detail = parity(wl, full.cpu().numpy(), res)
                par.update({"parity_full_size": ok, "parity_detail": detail})
                par["cpu_baseline"] = {"value": touched / cms / 1e3, "unit": "MTEPS", "cores": cores, "kind": kind,
                                       "sample": f"1 full bfs run from the bench source ({cms / 1e3:.1f} s); host "
                                                 f"cores available: {os.cpu_count()}", "parity_full_size": ok,
                                       "parity_detail": detail}
        except Exception as ex:
            par["error"] = str(ex)[:300]
        if touched is None:
            touched = total_edges
        value = touched * args.steps / ms / 1e3
        ach = 4.0 * inspected / world / (ms * 1e-3) / 1e9
        line = {"metric": f"MTEPS ({name})", "value": value, "unit": "MTEPS", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": wl["desc"],
                           "exchange": "kernels over NVLink peer memory (CUDA IPC windows)" if args.exchange == "p2p"
                           else "NCCL from the C++ level loop: grouped ncclSend/ncclRecv, ncclAllGather, ncclAllReduce",
                           "other_exchange": other_rec, "runs": runs_primary, "unsampled": unsampled,
                           "numerator": "sum of out-degrees of the reached vertices (identical at every N)",
                           "vertices": G.n_global, "edges": total_edges, "source": src,
                           "edges_touched_per_step": touched, "edges_inspected_per_step": inspected // args.steps,
                           "levels": st.levels, "level_direction": st.level_direction,
                           "level_frontier": st.level_frontier, "level_edges": st.level_edges,
                           "ids_exchanged_per_step_rank0": st.exchanged_ids,
                           "parity": {k: v for k, v in par.items() if k != "cpu_baseline"},
                           "l2_policy": "inputs larger than L2 per rank" if total_edges * 4 / world > 126e6 else
                                        "per-rank column indices %.0f MB" % (total_edges * 4 / world / 1e6)},
                "e2e": {"value": touched * args.steps / ms2 / 1e3, "unit": "MTEPS", "h2d_bytes_per_step": 4,
                        "d2h_bytes_per_step": G.n_local * 4, "ms_per_step": ms2 / args.steps},
                "gpu_launches": st.kernel_launches * args.steps,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             "traffic": None, "peak_kind": peak_kind,
                             "kernel": "whole step per rank (advance + sweep + exchange)", "bytes_per_edge": 4},
                "cpu_baseline": par.get("cpu_baseline"), "clocks": clocks}
        print(json.dumps(line), flush=True)
    mg.p2p_disconnect(eng, comm)
    G.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: bfs_do_rmat26 at N = 1, bfs_part_rmat26 (same graph, partitioned) at N > 1")
    ap.add_argument("--configs", default=None, choices=["all", "none"],
                    help="N = 1: also run the other BASELINE.json configurations as sub-records "
                         "(default: all when --workload is not given)")
    ap.add_argument("--scale", type=int, default=0, help="override the RMAT scale (testing only)")
    ap.add_argument("--lb", default=None, choices=["thread_mapped", "block_mapped", "merge_path"])
    ap.add_argument("--direction", default=None, choices=["forward", "backward", "optimized"])
    ap.add_argument("--hub-threshold", type=int, default=4096)
    ap.add_argument("--ctas-per-sm", type=int, default=8)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="partitioned workloads: frontier exchange by our kernels over peer memory, or NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reference-gpu", dest="reference_gpu", action="store_true", default=None,
                    help="also time the UNMODIFIED reference GPU kernels (oracle/_ref/gunrock_ref_gpu, built for "
                         "sm_100a with the atomics fix of SURVEY.md F2) on the same graphs and GPU; N = 1 only.  "
                         "Default: on for the default line (no --workload), off otherwise")
    ap.add_argument("--no-reference-gpu", dest="reference_gpu", action="store_false")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    explicit = args.workload is not None
    if args.reference_gpu is None:   # the same-GPU reference row belongs to the default line
        args.reference_gpu = (not explicit) and not args.scale and not args.no_cpu_baseline
    name = args.workload or (HEADLINE_N if world > 1 else HEADLINE_1)
    if args.impl == "reference" and not explicit:
        name = HEADLINE_N if args.gpus > 1 else HEADLINE_1
    wl = dict(WORKLOADS[name])

    def scaled(w):
        w = dict(w)
        if args.scale:
            w["scale"] = args.scale
            if w.get("pairs"):
                w["pairs"] = 17 * (1 << args.scale)
            if w["fold"]:
                w["fold"] = int(0.578 * (1 << args.scale))
        return w

    wl = scaled(wl)
    if args.lb:
        wl["lb"] = args.lb
    if args.direction:
        wl["direction"] = args.direction

    if args.impl == "reference":
        return run_reference(args, wl, name)

    import torch
    import torch.distributed as dist
    import gunrock_b200 as gb

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if gb.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libgunrock_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 or wl.get("partitioned"):
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), rank=rank, world_size=world)
    if wl.get("partitioned"):
        return run_partitioned(args, wl, name, rank, world, local)
    if world > 1:
        raise SystemExit(f"bench.py: workload {name} is a single-GPU configuration; N > 1 runs bfs_part_*")

    # ---- N = 1: the headline first (nothing else running), then the other configurations -------------
    sampler = ClockSampler(local)
    sampler.start()
    head, finish_head = bench_single(args, name, wl, args.steps, args.warmup, local,
                                     cpu_baseline=not args.no_cpu_baseline)
    clocks = sampler.stop()
    head["clocks"] = clocks
    run_configs = (args.configs or ("none" if explicit else "all")) == "all"
    subs = {}
    if run_configs:
        pending = []
        for sub in SUB_CONFIGS:
            if sub == name:
                continue
            try:
                k = min(args.steps, 10)
                rec, fin = bench_single(args, sub, scaled(WORKLOADS[sub]), k, args.warmup, local,
                                        cpu_baseline=not args.no_cpu_baseline,
                                        background_cpu=(sub == "sssp_rmat24"), extras=False)
                pending.append((sub, fin))
            except Exception as ex:
                subs[sub] = {"error": str(ex)[:400]}
        for sub, fin in pending:
            rec = fin()
            rec["config"].pop("runs", None)
            subs[sub] = rec
        head["configs"] = subs
    print(json.dumps(finish_head()), flush=True)


if __name__ == "__main__":
    main()
