#!/bin/bash
# Round 2, GPU call P (1 GPU): the whole GPU suite on the final code, the default bench line, ncu --set full of the
# sub-records' dominant kernels (SSSP RMAT-24 block_mapped: binned + hub; PageRank: tile kernel) for roofline.traffic.
set -u
OUT=gpurun_out/r2p
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > "$OUT/pytest_gpu.txt"
tail -4 "$OUT/pytest_gpu.txt"
( time python bench.py --steps 20 --warmup 5 ) > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
tail -4 "$OUT/bench_default.err"
python -c "
import json
j=json.load(open('$OUT/bench_default.json')); r=j['roofline']
print('ms/step %.3f e2e %.3f' % (j['ms_per_step'], j['e2e']['ms_per_step']), 'frac %.4f' % r['frac'], 'parity', j['cpu_baseline'].get('parity_full_size'))
for k,v in j['configs'].items(): print(k, '%.3f ms' % v['ms_per_step'], 'frac %.3f' % v['roofline']['frac'], 'parity', v['cpu_baseline'].get('parity_full_size'))
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'advance_binned_kernel|advance_hub_kernel' --launch-skip 48 --launch-count 16 -f \
  -o "$OUT/ncu_sssp24_block" python bench.py --workload sssp_rmat24 --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_sssp.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_sssp24_block.ncu-rep" "$OUT/ncu_sssp24_block.md" > /dev/null 2>&1 || true
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pr_pull_tile_kernel' --launch-skip 18 --launch-count 2 -f \
  -o "$OUT/ncu_pr_lj_tile" python bench.py --workload pr_lj --steps 1 --warmup 3 --no-cpu-baseline > "$OUT/ncu_pr.log" 2>&1
python profiles/summarize_ncu.py "$OUT/ncu_pr_lj_tile.ncu-rep" "$OUT/ncu_pr_lj_tile.md" > /dev/null 2>&1 || true
rm -f "$OUT"/ncu_sssp24_block.ncu-rep   # 16 full captures: too large to bring back; the summary stays
ls -la "$OUT"
