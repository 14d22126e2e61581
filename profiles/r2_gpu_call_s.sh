#!/bin/bash
# Round 2, GPU call S (1 GPU): final validation -- build check, smoke, the whole GPU suite, the default bench line and
# the reference arm exactly as the driver runs them (--steps 20 --warmup 5).
set -u
OUT=gpurun_out/r2s
mkdir -p "$OUT"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > "$OUT/smoke.txt" 2>&1; tail -2 "$OUT/smoke.txt"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"
tail -3 "$OUT/pytest_gpu.txt"
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"
python -c "
import json
j=json.load(open('$OUT/bench_default.json')); r=j['roofline']
print('ms/step %.3f e2e %.3f' % (j['ms_per_step'], j['e2e']['ms_per_step']), 'frac %.4f' % r['frac'], 'ceiling frac %.3f' % r['metric_ceiling']['value_frac'], 'parity', j['cpu_baseline'].get('parity_full_size'), 'launches', j['gpu_launches'], 'clocks', j['clocks'])
for k,v in j['configs'].items(): print(k, '%.3f ms' % v['ms_per_step'], 'frac %.3f' % v['roofline']['frac'], 'traffic', v['roofline']['traffic'], 'parity', v['cpu_baseline'].get('parity_full_size'))
"
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
tail -4 "$OUT/bench_reference.err"; cut -c1-400 "$OUT/bench_reference.json"
ls -la "$OUT"
