#!/bin/bash
# Round 2, GPU call H (1 GPU): full GPU suite on the round's final code, the default bench line (with the same-GPU
# reference row), the reference arm, launch list + ncu --set full of the headline's pull kernels and of the sub-configs'
# dominant kernels.
set -u
OUT=gpurun_out/r2h
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
( time python bench.py --steps 20 --warmup 5 --reference-gpu ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"; cut -c1-400 "$OUT/bench_default.json"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
tail -3 "$OUT/bench_reference.err"; cut -c1-300 "$OUT/bench_reference.json"
B2G_NO_PACKED_D2H=1 python bench.py --workload bfs_do_rmat26 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/do26_plain_d2h.json"
python -c "
import json
for f in ('bench_default','do26_plain_d2h'):
    j=json.load(open('$OUT/'+f+'.json')); print(f, 'ms/step %.3f e2e ms/step %.3f' % (j['ms_per_step'], j['e2e']['ms_per_step']), j['config']['level_kernel_ms'])
"
# launch list of one default-workload run (kernel shares), then --set full of the pull kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_do26.csv" \
  python bench.py --workload bfs_do_rmat26 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/launches_do26.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bfs_pull' --launch-skip 6 --launch-count 4 -f \
  -o "$OUT/ncu_bfs_pull_rmat26" python bench.py --workload bfs_do_rmat26 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_pull.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_bfs_pull_rmat26.ncu-rep" "$OUT/ncu_bfs_pull_rmat26.md" > /dev/null 2>&1 || true
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'advance_warp_path_kernel' --launch-skip 6 --launch-count 2 -f \
  -o "$OUT/ncu_push22_warp_path" python bench.py --workload bfs_push_rmat22 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_push.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_push22_warp_path.ncu-rep" "$OUT/ncu_push22_warp_path.md" > /dev/null 2>&1 || true
ls -la "$OUT"
