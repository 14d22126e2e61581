#!/bin/bash
# Round 2, GPU call C (1 GPU): sweep-only pull kernels (K1 + K2 over words), promoted warp-path merge_path, slab-table
# hub bin.  Full GPU suite, per-workload A/B, the default line, ncu of the pull and hub kernels.
set -u
OUT=gpurun_out/r2c
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "value %.0f" % j["value"], "level ms", c["level_kernel_ms"][:8],
          "roofline %.4f" % j["roofline"]["frac"], "outside %.3f" % c["outside_kernels_frac"])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
}
run() { # name, env..., -- bench args
  local name=$1; shift
  env "$@" 2>&1 | tail -1 > "$OUT/$name.json"; show "$OUT/$name.json"
}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline"
run do26_new   $B --workload bfs_do_rmat26
run do26_legacy B2G_BFS_PULL_LEGACY=1 $B --workload bfs_do_rmat26
run do22_new   $B --workload bfs_do_rmat22
run push22_v1  $B --workload bfs_push_rmat22
run push22_v0  B2G_ADVANCE_VARIANT=0 $B --workload bfs_push_rmat22
run push22_block $B --workload bfs_push_rmat22 --lb block_mapped
run sssp24_block $B --workload sssp_rmat24
run sssp24_block_h1024 $B --workload sssp_rmat24 --hub-threshold 1024
run sssp24_block_h65536 $B --workload sssp_rmat24 --hub-threshold 65536
run sssp24_merge $B --workload sssp_rmat24 --lb merge_path
run sssp24_merge_v0 B2G_ADVANCE_VARIANT=0 $B --workload sssp_rmat24 --lb merge_path
run pr_lj $B --workload pr_lj
( time python bench.py --steps 10 --warmup 3 --reference-gpu ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-300 "$OUT/bench_default.json"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bfs_pull' --launch-skip 6 --launch-count 4 -f \
  -o "$OUT/ncu_bfs_pull_rmat26" python bench.py --workload bfs_do_rmat26 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_pull.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_bfs_pull_rmat26.ncu-rep" "$OUT/ncu_bfs_pull_rmat26.md" > /dev/null 2>&1 || true
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'advance_(hub|binned)_kernel' --launch-skip 12 --launch-count 4 -f \
  -o "$OUT/ncu_sssp_block" python bench.py --workload sssp_rmat24 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_sssp.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_sssp_block.ncu-rep" "$OUT/ncu_sssp_block.md" > /dev/null 2>&1 || true
ls -la "$OUT"
