#!/bin/bash
# Round 2, GPU call E (1 GPU): K1 walks a queue of the unvisited bits and probes two stored in-neighbours (head[]),
# K2 starts at the third; out-degree sum at the pull -> push switch so that the tail kernel takes the rest.
set -u
OUT=gpurun_out/r2e
mkdir -p "$OUT"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_switches.py tests/test_gpu_multi.py tests/test_gpu_examples.py tests/test_pygunrock.py -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]; r = c["runs"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "runs best/med/worst %.3f %.3f %.3f" % (r["best_ms"], r["median_ms"], r["worst_ms"]),
          "level ms", c["level_kernel_ms"][:8], "edges", c["level_edges"][:8], "roofline %.4f" % j["roofline"]["frac"], "outside %.3f" % c["outside_kernels_frac"])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
}
run() { local name=$1; shift; env "$@" 2>&1 | tail -1 > "$OUT/$name.json"; show "$OUT/$name.json"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run do26 $B --workload bfs_do_rmat26
run do26_legacy B2G_BFS_PULL_LEGACY=1 $B --workload bfs_do_rmat26
run do22 $B --workload bfs_do_rmat22
run do26_pull $B --workload bfs_do_rmat26 --direction backward
( time python bench.py --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-300 "$OUT/bench_default.json"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bfs_pull' --launch-skip 6 --launch-count 4 -f \
  -o "$OUT/ncu_bfs_pull_rmat26" python bench.py --workload bfs_do_rmat26 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_pull.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_bfs_pull_rmat26.ncu-rep" "$OUT/ncu_bfs_pull_rmat26.md" > /dev/null 2>&1 || true
ls -la "$OUT"
