#!/bin/bash
# Round 2, GPU call N (2 GPUs): the multi-GPU test file (torchrun worker with the C++ NCCL loops of SSSP / PageRank,
# weighted partitioned PageRank), the N = 2 bench line on the final kernels, its phase trace, and the NVLink data
# counters of GPU 0 around 200 peer-memory traversals.
set -u
OUT=gpurun_out/r2n
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -25 > "$OUT/pytest_multi.txt"
tail -4 "$OUT/pytest_multi.txt"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580"
( time $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
tail -1 "$OUT/bench_n2.json" | cut -c1-600
( B2G_TRACE=1 $TR2 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline --exchange p2p ) > "$OUT/trace_n2.json" 2> "$OUT/trace_n2.err"
grep "b2g-p2p\] rank 0 phases" "$OUT/trace_n2.err" | tail -1 | cut -c1-700
nvidia-smi nvlink -gt d -i 0 > "$OUT/nvlink_before_p2p.txt" 2>&1
$TR2 bench.py --gpus 2 --steps 200 --warmup 3 --no-cpu-baseline --exchange p2p > "$OUT/nvlink_run_p2p.json" 2> /dev/null
nvidia-smi nvlink -gt d -i 0 > "$OUT/nvlink_after_p2p.txt" 2>&1
python - "$OUT" <<'PY'
import re, sys, json
out = sys.argv[1]
def total(f):
    tx = rx = 0
    for l in open(f):
        m = re.search(r"Data Tx:\s*(\d+)\s*KiB", l)
        if m: tx += int(m.group(1))
        m = re.search(r"Data Rx:\s*(\d+)\s*KiB", l)
        if m: rx += int(m.group(1))
    return tx, rx
try:
    a, b = total(f"{out}/nvlink_before_p2p.txt"), total(f"{out}/nvlink_after_p2p.txt")
    print("GPU 0 NVLink over one bench.py run of 203 peer-memory traversals + set-up + parity gather: tx %.1f MB rx %.1f MB" % ((b[0]-a[0])/1024, (b[1]-a[1])/1024))
except Exception as ex:
    print("nvlink counters unavailable:", ex, open(f"{out}/nvlink_after_p2p.txt").read()[:300])
PY
ls -la "$OUT"
