#!/bin/bash
# What was left unmeasured when round 1's GPU budget ran out, as 1-GPU gpurun calls.  Sections run in the order of
# their value; pick some with SECTIONS (default: all, about 25-30 GPU-minutes):
#   /usr/local/graft/bin/gpurun --timeout 900  -- 'SECTIONS="variants" bash profiles/next_round_first_gpu_call.sh'
#   /usr/local/graft/bin/gpurun --timeout 2700 -- 'bash profiles/next_round_first_gpu_call.sh'
# Results land in gpurun_out/next_round/.  Nothing printed under a profiler is a bench value.
set -u
OUT=gpurun_out/next_round
mkdir -p "$OUT"
SECTIONS=${SECTIONS:-all}
want() { [[ "$SECTIONS" == all || " $SECTIONS " == *" $1 "* ]]; }

# variants: the experimental merge_path kernels (advance.cuh advance_launch_t::variant) on the headline workload:
#   0 default | 1 warp-private spans (6 CTAs/SM) | 4 same, 8 chunks in flight | 7 = 4 + row-window prefetch
#   2 warp-private + on-chip visited copy (one CTA) | 5 / 6 copy over a cluster of 2 / 4 CTAs (DSMEM) | 3 4096-edge tiles
# A variant's number counts only once its bit-exactness tests (section "tests") pass; a first sanity check is
# printed here: edges touched (the out-degree sum of the reached vertices) must not depend on the variant.
if want variants; then
  for v in 0 1 4 7 2 5 6 3; do
    B2G_ADVANCE_VARIANT=$v python bench.py --workload bfs_push_rmat22 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/bfs_push_variant_$v.json"
    python - "$OUT" $v <<'PY'
import json, sys
out, v = sys.argv[1], sys.argv[2]
try:
    j = json.loads(open(f"{out}/bfs_push_variant_{v}.json").read())
    print(f"variant {v} bfs_push: {j['value']:.0f} MTEPS, {j['ms_per_step']:.3f} ms/step, roofline {j['roofline']['frac']:.3f}, "
          f"level ms {j['config']['level_kernel_ms'][:6]}, edges touched {j['config']['edges_touched_per_step']} "
          f"(must equal variant 0's), levels {j['config']['levels']}")
except Exception as ex:
    print(f"variant {v} bfs_push: no line ({ex})")
PY
  done | tee "$OUT/variants.txt"
fi

# tests: the whole GPU suite (incl. the never-run tests of tests/test_zz_gpu_widening.py and the pygunrock surface),
# then the opt-in bit-exactness tests of every experimental path
if want tests; then
  timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_gpu.txt"
  B2G_RUN_EXPERIMENTAL=1 timeout 1500 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_experimental.txt"
fi

# sssp: near/far schedule and the kernel variants on SSSP RMAT-24 (merge_path)
if want sssp; then
  for d in "" 8 16; do
    B2G_SSSP_DELTA=$d python bench.py --workload sssp_rmat24 --lb merge_path --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/sssp_delta_${d:-off}.json"
  done
  for v in 1 4 7 3; do
    B2G_ADVANCE_VARIANT=$v python bench.py --workload sssp_rmat24 --lb merge_path --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/sssp_variant_$v.json"
  done
  # configs[2] proper (block_mapped): which rows should go to the TMA hub bin?
  for h in 1024 4096 16384 65536; do
    python bench.py --workload sssp_rmat24 --hub-threshold $h --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/sssp_block_mapped_hub_$h.json"
  done
  python - "$OUT" <<'PY' | tee "$OUT/sssp.txt"
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/sssp_*.json")):
    try:
        j = json.loads(open(f).read())
        print(f"{f.split('/')[-1]}: {j['value']:.0f} MTEPS, {j['ms_per_step']:.3f} ms/step, {j['config']['levels']} iterations")
    except Exception as ex:
        print(f"{f}: no line ({ex})")
PY
fi

# do: direction-optimised BFS (configs[4] at one GPU): default pull levels vs the unvisited list from the first one
if want do; then
  for wl in bfs_do_rmat22 bfs_do_rmat26; do
    python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/${wl}_default.json"
    B2G_BFS_PULL_LIST_FIRST=1 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/${wl}_list_first.json"
  done
  python - "$OUT" <<'PY' | tee "$OUT/do.txt"
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bfs_do_*.json")):
    try:
        j = json.loads(open(f).read())
        print(f"{f.split('/')[-1]}: {j['ms_per_step']:.3f} ms/step, level ms {j['config']['level_kernel_ms'][:8]}, directions {j['config']['level_direction'][:8]}")
    except Exception as ex:
        print(f"{f}: no line ({ex})")
PY
fi

# refgpu: same-GPU baseline, the UNMODIFIED reference GPU kernels (oracle/_ref/gunrock_ref_gpu) on the bench graph
if want refgpu; then
  python bench.py --workload bfs_push_rmat22 --steps 5 --warmup 3 --no-cpu-baseline --reference-gpu 2>&1 | tail -1 > "$OUT/bench_with_reference_gpu.json"
  python -c "import json,sys; j=json.load(open('$OUT/bench_with_reference_gpu.json')); print('ours', round(j['value']), 'MTEPS; reference GPU:', json.dumps(j.get('reference_gpu'))[:600])" | tee "$OUT/reference_gpu.txt"
fi

# ncu: --set full of the two big levels (launches after the warm-up runs) for the default kernel, the barrier-free
# kernel and the whole-map cluster copy; summaries go next to the reports (copy the interesting ones into profiles/)
if want ncu; then
  for v in 0 4 6; do
    B2G_ADVANCE_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on \
      -k regex:'advance_(merge|warp)_path_kernel' --launch-skip 6 --launch-count 2 -f -o "$OUT/ncu_bfs_push_variant_$v" \
      python bench.py --workload bfs_push_rmat22 --steps 2 --warmup 3 --no-cpu-baseline > "$OUT/ncu_variant_$v.log" 2>&1
    python profiles/summarize_ncu.py "$OUT/ncu_bfs_push_variant_$v.ncu-rep" "$OUT/ncu_bfs_push_variant_$v.md" >/dev/null 2>&1 || true
  done
fi

# micro: probe rates of L1 / L2 / shared / DSMEM
if want micro; then
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o "$OUT/probe_rates" profiles/micro/probe_rates.cu && "$OUT/probe_rates" | tee "$OUT/probe_rates.txt"
fi

# bench: the default line, CPU baseline included
if want bench; then
  python bench.py --steps 10 --warmup 3 2>&1 | tail -1 > "$OUT/bench_default.json"
  cut -c1-300 "$OUT/bench_default.json"
fi
