#!/bin/bash
# Round 2, GPU call B (1 GPU): full GPU suite on the new code, headline A/B (second-generation pull kernels vs
# the first), the default bench line with every sub-record, the reference arm, the opt-in variant tests, ncu of
# the pull kernels.  Results under gpurun_out/r2b/.  Nothing printed under a profiler is a bench value.
set -u
OUT=gpurun_out/r2b
mkdir -p "$OUT"
nproc > "$OUT/nproc.txt"; free -g | head -2 >> "$OUT/nproc.txt"; nvidia-smi -L >> "$OUT/nproc.txt"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
for mode in new legacy; do
  if [ $mode = legacy ]; then export B2G_BFS_PULL_LEGACY=1; else unset B2G_BFS_PULL_LEGACY; fi
  for wl in bfs_do_rmat26 bfs_do_rmat22; do
    python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/${wl}_${mode}.json"
    python - "$OUT/${wl}_${mode}.json" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[1])); c = j["config"]
    print(sys.argv[1].split("/")[-1], "%.3f ms/step" % j["ms_per_step"], "value %.0f" % j["value"], "level ms", c["level_kernel_ms"],
          "dirs", c["level_direction"], "edges", c["level_edges"], "roofline %.4f" % j["roofline"]["frac"])
except Exception as ex:
    print(sys.argv[1], "no line:", ex, open(sys.argv[1]).read()[-400:])
PY
  done
done
unset B2G_BFS_PULL_LEGACY
( time python bench.py --steps 10 --warmup 3 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-600 "$OUT/bench_default.json"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
tail -4 "$OUT/bench_reference.err"; cut -c1-400 "$OUT/bench_reference.json"
B2G_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -5 > "$OUT/pytest_experimental.txt"
tail -2 "$OUT/pytest_experimental.txt"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bfs_pull' --launch-skip 9 --launch-count 6 -f \
  -o "$OUT/ncu_bfs_pull_rmat26" python bench.py --workload bfs_do_rmat26 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_pull.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_bfs_pull_rmat26.ncu-rep" "$OUT/ncu_bfs_pull_rmat26.md" > /dev/null 2>&1 || true
ls -la "$OUT"
