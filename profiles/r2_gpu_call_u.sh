#!/bin/bash
# Round 2, GPU call U (2 GPUs): per-run spread of both exchanges at N = 2 (call T's NCCL comparison arm averaged
# 3.0 ms over 5 runs against 1.0 ms in every earlier call).
set -u
OUT=gpurun_out/r2u
mkdir -p "$OUT"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580"
$TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > "$OUT/bench_n2_p2p.json" 2> "$OUT/bench_n2_p2p.err"
$TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --exchange nccl > "$OUT/bench_n2_nccl.json" 2> "$OUT/bench_n2_nccl.err"
python - "$OUT" <<'PY'
import json, sys
for f in ("bench_n2_p2p", "bench_n2_nccl"):
    try:
        j = json.loads(open(f"{sys.argv[1]}/{f}.json").read().strip().splitlines()[-1]); c = j["config"]
        print(f, "ms/step %.3f" % j["ms_per_step"], "runs", c.get("runs"), "other", c["other_exchange"])
    except Exception as ex:
        print(f, "no line:", ex, open(f"{sys.argv[1]}/{f}.err").read()[-600:])
PY
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1; head -6 "$OUT/topo.txt"
