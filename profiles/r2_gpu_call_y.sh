#!/bin/bash
# Round 2, GPU call Y (2 GPUs): call X again with the sampler child waited for before the warm-up steps.
set -u
OUT=gpurun_out/r2y
mkdir -p "$OUT"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580"
timeout 200 $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
echo "rc=$?"
python - "$OUT" <<'PY'
import json, sys
try:
    j = json.loads(open(f"{sys.argv[1]}/bench_n2.json").read().strip().splitlines()[-1]); c = j["config"]
    print("ms/step %.3f" % j["ms_per_step"], "e2e %.3f" % j["e2e"]["ms_per_step"], "\nruns", c.get("runs"), "\nunsampled", c.get("unsampled"), "\nother", c["other_exchange"], "\nclocks", j["clocks"], "\nparity", c["parity"])
except Exception as ex:
    print("no line:", ex, open(f"{sys.argv[1]}/bench_n2.err").read()[-1500:])
PY
