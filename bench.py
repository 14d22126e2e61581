#!/usr/bin/env python
"""bench.py -- MTEPS of the frontier hot path on synthetic RMAT graphs (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W            # our CUDA path (one JSON line)
    python bench.py --impl reference --gpus 1 --steps K ...  # the reference's CPU path, same config

A "step" is one full run of the workload's algorithm (one BFS / SSSP from the bench source, one
PageRank solve) over the graph already resident in HBM.  Default workload at N=1 is BASELINE.json
configs[1]: BFS push on RMAT-22 (ef 16), merge_path advance + in-kernel compact filter.
  value   : MTEPS, device-resident result, CUDA events on the launching stream, max over ranks
  e2e     : same metric through the C-ABI call with HOST buffers (source in, distances out to
            pinned host memory inside the timed region)
  roofline: dominant kernel's algorithmic bytes (4 B x column indices read; 8 B for SSSP) divided by
            its device time (per-level CUDA events recorded on the launching stream by the enactor)
  cpu_baseline: the reference's own CPU validator (oracle/_ref, compiled from /root/reference) or,
            when that binary is absent, the oracle's C port, timed on this box's host cores.
N > 1 (torchrun): every rank holds the same RMAT graph and runs BFS from its own shard of a batch of
sources (independent objects, no data-path collective) -> "scaling": "weak".
"""
import argparse
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
    # name: (algorithm, scale, pairs-per-vertex, seed, mirror, fold, weights, lb, direction)
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
    # config 5: 1-D vertex partition over the ranks + NCCL frontier exchange (strong scaling)
    "bfs_part_rmat26": dict(alg="bfs", scale=26, ef=8, seed=0x5EED26, mirror=True, fold=0, weights=0,
                            lb="block_mapped", direction="optimized", partitioned=True,
                            desc="BFS direction-optimised, RMAT-26 (16 directed edges/vertex), 1-D cyclic vertex "
                                 "partition across the ranks, NCCL all-to-all / all-gather frontier exchange"),
    "bfs_part_rmat22": dict(alg="bfs", scale=22, ef=16, seed=0x5EED22, mirror=True, fold=0, weights=0,
                            lb="block_mapped", direction="optimized", partitioned=True,
                            desc="BFS direction-optimised, RMAT-22 ef16, 1-D partition + NCCL exchange"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
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
                "samples": len(sm), "reasons": sorted(reasons)}


def host_graph(wl):
    """The workload's CSR built on the HOST by the checker's generator (reference arm / CPU baseline
    when no device copy is at hand).  Bit-identical to the device generator (tests/test_gpu_parity)."""
    import oracle
    V = wl["fold"] or (1 << wl["scale"])
    n_pairs = wl.get("pairs") or wl["ef"] * (1 << wl["scale"])
    s, d = oracle.rmat_edges(wl["scale"], n_pairs, wl["seed"])
    if wl["fold"]:
        s, d = (s % wl["fold"]).astype(np.int32), (d % wl["fold"]).astype(np.int32)
    ro, ci = oracle.build_csr_from_pairs(V, s, d, wl["mirror"])
    w = oracle.edge_weights(wl["seed"] + 1, ro, ci, wl["weights"] == 2) if wl["weights"] else None
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


def edges_touched_cpu(wl, ro, result):
    deg = np.diff(ro).astype(np.int64)
    if wl["alg"] == "bfs":
        return int(deg[result < 2**31 - 1].sum())
    if wl["alg"] == "sssp":
        return int(deg[result < np.finfo(np.float32).max].sum())  # each reached vertex popped >= once
    p, iters = result
    return int(deg.sum()) * int(iters)


def bench_source(ro):
    deg = np.diff(ro)
    return int(deg.argmax())


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
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
            if r.returncode != 0:
                out[lb] = {"error": (r.stderr or r.stdout)[-300:]}
                continue
            j = json.loads(r.stdout.strip().splitlines()[-1])
            ms = sorted(j["ms"][1:] or j["ms"])   # first run pays the lazy module load
            out[lb] = {"ms_median": statistics.median(ms), "ms_best": ms[0], "runs": len(j["ms"]),
                       "errors_vs_reference_cpu": j["errors"],
                       "mteps": (edges_per_run / statistics.median(ms) / 1e3) if wl["alg"] != "pr" else None}
    return {"kind": "unmodified reference GPU kernels, nvcc sm_100a, -include oracle/ref_gpu_fix.h, SM_TARGET=90",
            "numerator": "edges touched by OUR run of the same traversal (bfs / sssp); pr: time per solve only",
            **out}


def run_reference(args, wl, name):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.time()
    ro, ci, w = host_graph(wl)
    run, kind, cores = cpu_run_factory(wl, ro, ci, w)
    src = bench_source(ro)
    for _ in range(args.warmup):
        run(src)
    ms, res = [], None
    for _ in range(args.steps):
        res, t = run(src)
        ms.append(t)
    et = edges_touched_cpu(wl, ro, res)
    total_ms = sum(ms)
    value = et * len(ms) / total_ms / 1e3
    line = {"impl": "reference", "metric": f"MTEPS ({name})", "value": value, "unit": "MTEPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(len(ms), 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32" if wl["alg"] == "bfs" else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "vertices": int(len(ro) - 1), "edges": int(len(ci)),
                       "source": src, "edges_touched_per_step": et},
            "cpu_baseline": {"value": value, "unit": "MTEPS", "cores": cores, "kind": kind,
                             "sample": f"{len(ms)} full {wl['alg']} run(s) from the bench source, "
                                       f"validator's own timer; host cores available: {os.cpu_count()}"},
            "e2e": {"value": value, "unit": "MTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "setup_s": round(time.time() - t0 - total_ms / 1e3, 1)}
    print(json.dumps(line), flush=True)


def run_partitioned(args, wl, name, rank, world, local):
    """BASELINE.json configs[4]: one BFS over a graph 1-D partitioned across the ranks (strong scaling)."""
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

    exchange = args.exchange
    if exchange == "p2p":     # the kernels exchange frontiers over NVLink peer memory (bfs_p2p.cuh)
        mg.p2p_connect(eng, comm)

        def step():
            return mg.bfs_rank_p2p(eng, src, total_edges, direction)
    else:                     # NCCL all-to-all / all-gather / all-reduce between the per-rank kernels
        def step():
            return mg.bfs_rank_async(eng, comm, src, total_edges, direction=direction)

    for _ in range(max(args.warmup, 3)):
        dloc, st = step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    edges = 0
    for _ in range(args.steps):
        dloc, st = step()
        edges += st.edges_touched
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # e2e: the rank's slice of the result is copied to pinned host memory inside the timed region
    h = torch.empty(G.n_local, dtype=torch.int32).pin_memory()
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        dloc, st2 = step()
        h.copy_(dloc, non_blocking=False)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms2 = float(ms2.item())
    reached = torch.tensor([int((dloc < 2**31 - 1).sum())], device="cuda")
    dist.all_reduce(reached)
    agree = None
    if exchange == "p2p":     # outside the timed region: same depths as the NCCL-exchange path, on every rank
        d_nccl, _ = mg.bfs_rank_async(eng, comm, src, total_edges, direction=direction)
        same = torch.tensor([int(torch.equal(dloc, d_nccl))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        agree = bool(same.item())
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        peak, peak_kind = peaks()
        value = edges / ms / 1e3
        line = {"metric": f"MTEPS ({name})", "value": value, "unit": "MTEPS", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": {"workload": wl["desc"], "exchange": "kernels over NVLink peer memory (CUDA IPC windows)"
                           if exchange == "p2p" else "NCCL all_to_all_single / all_gather / all_reduce",
                           "vertices": G.n_global, "edges": total_edges, "source": src,
                           "levels": st.levels, "level_direction": st.level_direction,
                           "level_frontier": st.level_frontier, "level_edges": st.level_edges,
                           "ids_exchanged_per_step_rank0": st.exchanged_ids, "reached_vertices": int(reached.item()),
                           "depths_equal_nccl_path": agree,
                           "l2_policy": "inputs larger than L2 per rank" if total_edges * 4 / world > 126e6 else
                                        "per-rank column indices %.0f MB" % (total_edges * 4 / world / 1e6),
                           "graph500_mteps": total_edges / (ms / args.steps) / 1e3},
                "e2e": {"value": edges / ms2 / 1e3, "unit": "MTEPS", "h2d_bytes_per_step": 4,
                        "d2h_bytes_per_step": G.n_local * 4, "ms_per_step": ms2 / args.steps},
                "gpu_launches": (st.kernel_launches * args.steps) if exchange == "p2p" else None,
                "roofline": {"bound": "hbm", "achieved": 4.0 * edges / world / (ms * 1e-3) / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": 4.0 * edges / world / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                             "peak_kind": peak_kind, "kernel": "whole step per rank (advance + sweep + exchange)",
                             "bytes_per_edge": 4},
                "cpu_baseline": None, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if exchange == "p2p":
        mg.p2p_disconnect(eng, comm)
    G.close()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bfs_push_rmat22", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=int, default=0, help="override the RMAT scale (testing only)")
    ap.add_argument("--lb", default=None, choices=["thread_mapped", "block_mapped", "merge_path"])
    ap.add_argument("--direction", default=None, choices=["forward", "backward", "optimized"])
    ap.add_argument("--hub-threshold", type=int, default=4096)
    ap.add_argument("--ctas-per-sm", type=int, default=8)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="partitioned workloads: frontier exchange by our kernels over peer memory, or NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reference-gpu", action="store_true",
                    help="also time the UNMODIFIED reference GPU kernels (oracle/_ref/gunrock_ref_gpu, built for "
                         "sm_100a with the atomics fix of SURVEY.md F2) on the same graph and GPU; N = 1 only")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    name = args.workload
    wl = dict(WORKLOADS[name])
    if args.scale:
        wl["scale"] = args.scale
        if wl.get("pairs"):
            wl["pairs"] = 17 * (1 << args.scale)
        if wl["fold"]:
            wl["fold"] = int(0.578 * (1 << args.scale))
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
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

    # ---- graph, generated on the device (ingest is untimed, as in the reference) ----------------
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

    # N > 1: a batch of sources, sharded over ranks (replicated graph, no collective on the data path)
    # The unit of work is one traversal from a hub: N = 1 runs the highest-degree vertex (the bench source
    # of the reference arm and of the CPU baseline); N > 1 runs the 4 N highest-degree vertices, 4 per
    # rank, so that every rank's traversals have the same level profile as the N = 1 unit.
    sources = [src]
    if world > 1:
        deg_host = np.diff(G.download()[0])
        batch = [int(x) for x in np.argsort(-deg_host.astype(np.int64), kind="stable")[:world * 4]]
        sources = batch[rank::world]

    def step(out):
        tot = None
        for s in (sources if world > 1 else [src]):
            if wl["alg"] == "bfs":
                st = gb.bfs(G, s, out, options=opt)
            elif wl["alg"] == "sssp":
                st = gb.sssp(G, s, out, options=opt)
            else:
                st = gb.pr(G, out, 0.85, 1e-6, options=opt)
            if tot is None:
                tot = st
            else:
                tot.edges_touched += st.edges_touched
                tot.kernel_launches += st.kernel_launches
                tot.level_edges += st.level_edges
                tot.level_kernel_ms += st.level_kernel_ms
        return tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(out, steps):
        """K steps bracketed by barrier+sync; CUDA events on the launching stream; max over ranks."""
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        agg = dict(edges=0, launches=0, kern_bytes=0.0, kern_ms=0.0, kern_launches=0, run_ms=[])
        barrier()
        t_wall = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                st = step(out)
                agg["edges"] += st.edges_touched
                agg["launches"] += st.kernel_launches
                agg["run_ms"].append(float(st.elapsed_ms))   # the library's own events around enact() (first source)
                for e, ms in zip(st.level_edges, st.level_kernel_ms):
                    agg["kern_bytes"] += e * bytes_per_edge
                    agg["kern_ms"] += ms
                    agg["kern_launches"] += 1
            e1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t_wall) * 1e3
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, float(agg["edges"]), float(agg["launches"])], device="cuda", dtype=torch.float64)
            mx = t.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, agg["edges"], agg["launches"] = float(mx[0]), int(t[1]), int(t[2])
        return ms, wall_ms, agg, st

    bytes_per_edge = 8 if wl["alg"] == "sssp" else 4
    for _ in range(args.warmup):
        step(d_out)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, wall_ms, agg, last = timed(d_out, args.steps)
    for _ in range(1):
        step(h_out)
    ms_e2e, wall_e2e, agg_e2e, _ = timed(h_out, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    value = agg["edges"] / ms / 1e3                       # edges / ms / 1000 = MTEPS (performance.hxx:225-229)
    e2e = agg_e2e["edges"] / ms_e2e / 1e3
    peak, peak_kind = peaks()
    # roofline for the dominant kernel (this rank): algorithmic bytes per launch / mean launch time
    ach = agg["kern_bytes"] / (agg["kern_ms"] * 1e-3) / 1e9 if agg["kern_ms"] > 0 else 0.0
    if world > 1 and rank != 0:
        dist.destroy_process_group()
        return

    # DRAM traffic per launch of the dominant kernel: from the committed ncu capture of this very
    # configuration (profiles/r1_traffic.json), never measured under the profiler here
    lbk = "merge_path" if wl["lb"] == "merge_path" else "binned"
    traffic, traffic_detail = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            entry = json.load(f).get(f"{name}/{wl['lb']}")
        if entry and wl["direction"] == "forward" and not args.scale:
            per = [l["dram_read_bytes"] + l["dram_write_bytes"] for l in entry["launches"]]
            traffic = sum(per) / len(per)
            traffic_detail = {"source": entry["source"], "launches": entry["launches"],
                              "note": "mean over the capture's launches; compare with their algorithmic_bytes"}
    except (OSError, ValueError, KeyError):
        pass

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) -----------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            ro, ci, w = G.download()
            run, kind, cores = cpu_run_factory(wl, ro, ci, w)
            res, cms = run(src)
            et = edges_touched_cpu(wl, ro, res)
            cpu = {"value": et / cms / 1e3, "unit": "MTEPS", "cores": cores, "kind": kind,
                   "sample": f"1 full {wl['alg']} run from the bench source ({cms / 1e3:.1f} s, validator's own timer); "
                             f"host cores available: {os.cpu_count()}"}
            # parity spot check on the full-size result (the checker checking the product, not the reverse)
            if wl["alg"] == "pr":
                ok = bool(np.allclose(h_out.numpy(), res[0], rtol=1e-6, atol=0))
            else:
                ok = bool(np.array_equal(h_out.numpy().view(np.uint32), np.asarray(res).view(np.uint32)))
            cpu["parity_full_size"] = ok
        except Exception as ex:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "MTEPS", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}

    # Optional second baseline row: the reference's own GPU implementation on this GPU,
    # same graph (written once in the reference's .csr layout), same source, its own enactor timer.
    ref_gpu = None
    if args.reference_gpu and world == 1:
        try:
            ref_gpu = reference_gpu_leg(G, wl, src, agg["edges"] // args.steps)
        except Exception as ex:
            ref_gpu = {"error": str(ex)}

    # best / median of the runs (the reference's timed region: CUDA events around enact()),
    # and the mean over 16 random sources of degree > 0 (RNG seed 1).  Extras: they never take the line down.
    runs = None
    try:
        if world == 1 and agg["run_ms"]:
            r = sorted(agg["run_ms"])
            per_run_edges = agg["edges"] / args.steps
            runs = {"best_ms": r[0], "median_ms": statistics.median(r), "worst_ms": r[-1],
                    "best_mteps": per_run_edges / r[0] / 1e3, "median_mteps": per_run_edges / statistics.median(r) / 1e3,
                    "region": "CUDA events around the enactor loop of each run (enactor.hxx:266-288)"}
            if wl["alg"] in ("bfs", "sssp") and not args.no_cpu_baseline:
                deg_host = np.diff(G.download()[0])
                picks = np.random.default_rng(1).choice(np.flatnonzero(deg_host > 0), 16, replace=False)
                per = []
                with torch.cuda.stream(stream):
                    for s16 in picks:
                        fn = gb.bfs if wl["alg"] == "bfs" else gb.sssp
                        st16 = fn(G, int(s16), d_out, options=opt)
                        per.append(st16.edges_touched / max(st16.elapsed_ms, 1e-6) / 1e3)
                runs["random16_mean_mteps"] = float(np.mean(per))
                runs["random16_min_mteps"] = float(np.min(per))
    except Exception as ex:
        runs = {"error": str(ex)}

    out_bytes = V * 4
    line = {
        "metric": f"MTEPS ({name})", "value": value, "unit": "MTEPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32" if wl["alg"] == "bfs" else "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "vertices": V, "edges": G.n_edges, "source": src,
                   "source_degree": src_deg, "load_balance": wl["lb"], "direction": wl["direction"],
                   "filter": "compact (fused into advance)", "sources_per_step": len(sources) * world if world > 1 else 1,
                   "sources": "highest-degree vertex" if world == 1 else "the 4 N highest-degree vertices, 4 per rank",
                   "l2_policy": "inputs larger than L2 (column indices %.0f MB > 126 MB)" % (G.n_edges * 4 / 1e6),
                   "levels": last.iterations, "level_direction": last.level_direction,
                   "level_frontier": last.level_frontier, "level_edges": last.level_edges[:last.iterations],
                   "level_kernel_ms": [round(x, 4) for x in last.level_kernel_ms[:last.iterations]],
                   "edges_touched_per_step": agg["edges"] // args.steps, "runs": runs,
                   "experimental": {"B2G_ADVANCE_VARIANT": os.environ.get("B2G_ADVANCE_VARIANT", "0"),
                                    "B2G_SSSP_DELTA": os.environ.get("B2G_SSSP_DELTA", "")},
                   "graph500_mteps": (G.n_edges * (len(sources) * world if world > 1 else 1)) / (ms / args.steps) / 1e3},
        "e2e": {"value": e2e, "unit": "MTEPS", "h2d_bytes_per_step": 4 * (len(sources) if world > 1 else 1),
                "d2h_bytes_per_step": out_bytes * (len(sources) if world > 1 else 1), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": agg["launches"],
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                     "traffic": traffic, "traffic_detail": traffic_detail, "peak_kind": peak_kind,
                     "kernel": {"bfs": "advance_%s_kernel<bfs_claim_op>" % lbk,
                                "sssp": "advance_%s_kernel<sssp_relax_op>" % lbk, "pr": "pr_pull_tile_kernel"}[wl["alg"]],
                     "bytes_per_edge": bytes_per_edge, "launches": agg["kern_launches"],
                     "kernel_ms_total": agg["kern_ms"]},
        "cpu_baseline": cpu, "clocks": clocks, "wall_ms": wall_ms,
    }
    if ref_gpu is not None:
        line["reference_gpu"] = ref_gpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
