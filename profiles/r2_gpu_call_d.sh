#!/bin/bash
# Round 2, GPU call D (1 GPU): hub-bin ticket range fix (level 0), clock sampler A/B (in-process NVML vs none), the
# default line again, quick full-suite rerun of the files that changed.
set -u
OUT=gpurun_out/r2d
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_switches.py tests/test_gpu_multi.py tests/test_gpu_examples.py -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]; r = c["runs"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "runs best/med/worst %.3f %.3f %.3f" % (r["best_ms"], r["median_ms"], r["worst_ms"]),
          "level ms", c["level_kernel_ms"][:8], "roofline %.4f" % j["roofline"]["frac"], "clocks", j.get("clocks"))
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
}
run() { local name=$1; shift; env "$@" 2>&1 | tail -1 > "$OUT/$name.json"; show "$OUT/$name.json"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run do26_nvml $B --workload bfs_do_rmat26
run do26_nosampler B2G_BENCH_NO_SAMPLER=1 $B --workload bfs_do_rmat26
run do26_nvml_b $B --workload bfs_do_rmat26
run sssp24_block_nvml $B --workload sssp_rmat24
run sssp24_block_nosampler B2G_BENCH_NO_SAMPLER=1 $B --workload sssp_rmat24
run sssp24_merge $B --workload sssp_rmat24 --lb merge_path
run push22 $B --workload bfs_push_rmat22
run push22_block $B --workload bfs_push_rmat22 --lb block_mapped
run do22 $B --workload bfs_do_rmat22
( time python bench.py --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-300 "$OUT/bench_default.json"
ls -la "$OUT"
