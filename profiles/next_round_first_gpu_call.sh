#!/bin/bash
# What was left unmeasured when round 1's GPU budget ran out -- one 1-GPU gpurun call:
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash profiles/next_round_first_gpu_call.sh'   (about 25-30 GPU-minutes)
# Results land in gpurun_out/next_round/.  Nothing printed under a profiler is a bench value.
set -u
OUT=gpurun_out/next_round
mkdir -p "$OUT"
# 1. GPU tests written after the last GPU run (pygunrock surface, tc / spgemm / mst examples) + the whole suite
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
# 2. experimental near/far SSSP: bit-exactness first, then its time against the default schedule
B2G_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_experimental.txt"
for d in "" 8 16; do
  B2G_SSSP_DELTA=$d python bench.py --workload sssp_rmat24 --lb merge_path --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/sssp_delta_${d:-off}.json"
done
# 2b. experimental merge_path kernels (advance.cuh advance_launch_t::variant): default vs warp-private spans
#     (1: 6 CTAs/SM, 4: more loads in flight), warp-private + shared-memory visited snapshot (2; 5 / 6: spread over a
#     cluster of 2 / 4 CTAs through distributed shared memory), 4096-edge tiles (3)
for v in 0 1 2 3 4 5 6 7; do
  B2G_ADVANCE_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/bfs_push_variant_$v.json"
  B2G_ADVANCE_VARIANT=$v python bench.py --workload sssp_rmat24 --lb merge_path --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/sssp_variant_$v.json"
  python - "$OUT" $v <<'PY'
import json, sys
out, v = sys.argv[1], sys.argv[2]
for w in ("bfs_push", "sssp"):
    try:
        j = json.loads(open(f"{out}/{w}_variant_{v}.json").read())
        print(f"variant {v} {w}: {j['value']:.0f} MTEPS, {j['ms_per_step']:.3f} ms/step, roofline {j['roofline']['frac']:.3f}, "
              f"level ms {j['config']['level_kernel_ms'][:6]}")
    except Exception as ex:
        print(f"variant {v} {w}: no line ({ex})")
PY
done | tee "$OUT/variants.txt"
# 2c. same-GPU baseline: the UNMODIFIED reference GPU kernels (oracle/_ref/gunrock_ref_gpu) on the bench graph
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reference-gpu 2>&1 | tail -1 > "$OUT/bench_with_reference_gpu.json"
python -c "import json,sys; j=json.load(open('$OUT/bench_with_reference_gpu.json')); print('ours', round(j['value']), 'MTEPS; reference GPU:', json.dumps(j.get('reference_gpu'))[:600])" | tee "$OUT/reference_gpu.txt"
# 2d. ncu --set full of the two big levels (launches after the warm-up runs) for the default kernel and the two
#     on-chip-map variants; summaries go next to the reports (copy the interesting ones into profiles/)
for v in 0 2 6; do
  B2G_ADVANCE_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'advance_(merge|warp)_path_kernel' --launch-skip 6 --launch-count 2 -f -o "$OUT/ncu_bfs_push_variant_$v" \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_variant_$v.log" 2>&1
  python profiles/summarize_ncu.py "$OUT/ncu_bfs_push_variant_$v.ncu-rep" "$OUT/ncu_bfs_push_variant_$v.md" >/dev/null 2>&1 || true
done
# 3. design input for the on-chip visited map: probe rates of L1 / L2 / shared / DSMEM
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o "$OUT/probe_rates" profiles/micro/probe_rates.cu && "$OUT/probe_rates" | tee "$OUT/probe_rates.txt"
# 4. the default bench line (hub sources for N > 1 are measured by the 2-GPU call of the round)
python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > "$OUT/bench_default.json"
cut -c1-300 "$OUT/bench_default.json"
