#!/bin/bash
# Round 2, GPU call L (1 GPU, short): A/B of the pull kernels' batch size and of the lazy label fill on RMAT-22 / the
# per-rank share of RMAT-26 (scale 23 stands in for one of 8 ranks), TMA tail clamp re-check (parity tests).
set -u
OUT=gpurun_out/r2l
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_switches.py -m gpu -q 2>&1 | tail -3 > "$OUT/pytest_gpu.txt"
tail -2 "$OUT/pytest_gpu.txt"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]; r = c["runs"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "runs best/med/worst %.3f %.3f %.3f" % (r["best_ms"], r["median_ms"], r["worst_ms"]),
          "level ms", c["level_kernel_ms"][:8], "outside %.3f" % c["outside_kernels_frac"])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
}
run() { local name=$1; shift; env "$@" 2>&1 | tail -1 > "$OUT/$name.json"; show "$OUT/$name.json"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
for b in 8 16 32; do
  run do22_batch$b B2G_PULL_BATCH=$b $B --workload bfs_do_rmat22
  run do26s23_batch$b B2G_PULL_BATCH=$b $B --workload bfs_do_rmat26 --scale 23
done
run do22_eager_b32 B2G_PULL_BATCH=32 B2G_BFS_EAGER_FILL=1 $B --workload bfs_do_rmat22
run do26_batch16 B2G_PULL_BATCH=16 $B --workload bfs_do_rmat26
run do26_auto $B --workload bfs_do_rmat26
ls -la "$OUT"
