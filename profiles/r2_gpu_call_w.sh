#!/bin/bash
# Round 2, GPU call W (1 GPU, ~7 GPU-minutes left in the round): the new reference-output-layout selftest, the API
# selftests rebuilt against the changed headers, smoke(), and -- if the clock allows -- the headline alone.
set -u
OUT=gpurun_out/r2w
mkdir -p "$OUT"
for t in reference_layout_selftest api_selftest cache_selftest dense_frontier_selftest; do
  timeout 60 examples/bin/$t > "$OUT/$t.txt" 2>&1; echo "$t rc=$?" | tee -a "$OUT/summary.txt"
done
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.txt" 2>&1; echo "smoke rc=$?" | tee -a "$OUT/summary.txt"
tail -2 "$OUT/smoke.txt"
timeout 120 python -m pytest tests/test_zz_gpu_widening.py -q -m gpu -k "reference_output_layout or dense_frontier" > "$OUT/pytest_zz.txt" 2>&1
echo "pytest rc=$?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/pytest_zz.txt"
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --configs none > "$OUT/bench_headline.json" 2> "$OUT/bench_headline.err"
echo "bench rc=$?" | tee -a "$OUT/summary.txt"
python - "$OUT" <<'PY'
import json, sys
try:
    j = json.loads(open(f"{sys.argv[1]}/bench_headline.json").read().strip().splitlines()[-1])
    print("ms/step %.3f value %.0f e2e %.3f clocks %s" % (j["ms_per_step"], j["value"], j["e2e"]["ms_per_step"], j["clocks"]))
except Exception as ex:
    print("no bench line:", ex)
PY
