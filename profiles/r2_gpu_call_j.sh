#!/bin/bash
# Round 2, GPU call J (1 GPU): labels of unreached vertices written once per run (lazy fill), adaptive batch of the
# pull kernels.  Full suite + headline A/B + default line.
set -u
OUT=gpurun_out/r2j
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]; r = c["runs"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "e2e %.3f" % j["e2e"]["ms_per_step"], "runs best/med/worst %.3f %.3f %.3f" % (r["best_ms"], r["median_ms"], r["worst_ms"]),
          "level ms", c["level_kernel_ms"][:8], "roofline %.4f" % j["roofline"]["frac"], "outside %.3f" % c["outside_kernels_frac"])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
}
run() { local name=$1; shift; env "$@" 2>&1 | tail -1 > "$OUT/$name.json"; show "$OUT/$name.json"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run do26 $B --workload bfs_do_rmat26
run do26_eager B2G_BFS_EAGER_FILL=1 $B --workload bfs_do_rmat26
run do22 $B --workload bfs_do_rmat22
run do26_pullonly $B --workload bfs_do_rmat26 --direction backward
( time python bench.py --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-300 "$OUT/bench_default.json"
ls -la "$OUT"
