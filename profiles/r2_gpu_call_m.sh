#!/bin/bash
# Round 2, GPU call M (1 GPU): the new unpadded-view test with its traceback, full GPU suite, default line (now with
# the same-GPU reference row by default), and the TMA A/B on the SAME edges: BFS push RMAT-22 block_mapped with every
# row of >= 64 edges staged through cp.async.bulk slabs (hub threshold 64) against no row staged (threshold 2^30),
# bench lines + ncu --set full of both kernels.
set -u
OUT=gpurun_out/r2m
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k unpadded --tb=long 2>&1 | tail -80 > "$OUT/pytest_unpadded.txt"
tail -30 "$OUT/pytest_unpadded.txt"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
( time python bench.py --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-400 "$OUT/bench_default.json"
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload bfs_push_rmat22 --lb block_mapped"
for thr in 64 256 4096 1073741824; do
  $B --hub-threshold $thr 2>&1 | tail -1 > "$OUT/push22_block_thr$thr.json"
  python - "$OUT/push22_block_thr$thr.json" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "level ms", c["level_kernel_ms"][:8])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
done
for thr in 64 1073741824; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'advance_hub_kernel|advance_binned_kernel' --launch-skip 9 --launch-count 4 -f \
    -o "$OUT/ncu_push22_block_thr$thr" python bench.py --workload bfs_push_rmat22 --lb block_mapped --hub-threshold $thr --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_thr$thr.log" 2>&1
  python profiles/summarize_ncu.py "$OUT/ncu_push22_block_thr$thr.ncu-rep" "$OUT/ncu_push22_block_thr$thr.md" > /dev/null 2>&1 || true
done
ls -la "$OUT"
