#!/bin/bash
# Round 2, GPU call O (1 GPU): multi-device-context SSSP / PageRank with the ranks sharing the one device (thread
# exchange), the multi-GPU test file's single-rank tests, the packed D2H with one OpenMP team (thread-count sweep),
# and the default bench line with the new roofline objects.
set -u
OUT=gpurun_out/r2o
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_examples.py tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -30 > "$OUT/pytest_gpu.txt"
tail -5 "$OUT/pytest_gpu.txt"
for t in 8 16 24 48 96; do
  B2G_D2H_THREADS=$t python bench.py --workload bfs_do_rmat26 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/do26_d2h_threads$t.json"
  python -c "
import json
j=json.load(open('$OUT/do26_d2h_threads$t.json')); print('threads $t: ms/step %.3f e2e ms/step %.3f' % (j['ms_per_step'], j['e2e']['ms_per_step']))"
done
( time python bench.py --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"
python -c "
import json
j=json.load(open('$OUT/bench_default.json')); r=j['roofline']
print('ms/step %.3f e2e %.3f' % (j['ms_per_step'], j['e2e']['ms_per_step']), 'frac %.4f' % r['frac'], r.get('metric_ceiling'), r.get('pull_levels_full_traffic_model'))
for k,v in j['configs'].items(): print(k, '%.3f ms' % v['ms_per_step'], 'frac %.3f' % v['roofline']['frac'], v['roofline'].get('metric_ceiling',{}).get('value_frac'))
"
ls -la "$OUT"
