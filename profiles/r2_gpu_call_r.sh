#!/bin/bash
# Round 2, GPU call R (1 GPU): occupancy A/Bs of the latency-bound kernels.  PageRank tile kernel compiled for 6
# (default, 40 registers) against 8 resident CTAs per SM (32 registers, B2G_PR_CTAS=8); SSSP's binned + slab kernels
# in two A/B builds of the library (B2G_LIB_PATH): ab66 = both compiled for 6 CTAs per SM (40 registers), ab58 =
# slab kernel for 5 (48 registers), binned for 8 (32 registers, spills).
set -u
OUT=gpurun_out/r2r
mkdir -p "$OUT"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]; r = c["runs"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "runs best/med/worst %.3f %.3f %.3f" % (r["best_ms"], r["median_ms"], r["worst_ms"]),
          "level ms", c["level_kernel_ms"][:8])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
}
run() { local name=$1; shift; env "$@" 2>&1 | tail -1 > "$OUT/$name.json"; show "$OUT/$name.json"; }
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline"
run pr_default $B --workload pr_lj
run pr_ctas8 B2G_PR_CTAS=8 $B --workload pr_lj
run sssp_default $B --workload sssp_rmat24
run sssp_ab66 B2G_LIB_PATH=$PWD/gunrock_b200/libgunrock_b200_ab66.so $B --workload sssp_rmat24
run sssp_ab58 B2G_LIB_PATH=$PWD/gunrock_b200/libgunrock_b200_ab58.so $B --workload sssp_rmat24
run push22_block_default $B --workload bfs_push_rmat22 --lb block_mapped
run push22_block_ab66 B2G_LIB_PATH=$PWD/gunrock_b200/libgunrock_b200_ab66.so $B --workload bfs_push_rmat22 --lb block_mapped
run do26_ab66 B2G_LIB_PATH=$PWD/gunrock_b200/libgunrock_b200_ab66.so $B --workload bfs_do_rmat26
ls -la "$OUT"
