#!/bin/bash
# Round 2, GPU call K (4 GPUs): multi-GPU paths after the adaptive pull batches: torchrun worker (4 ranks), N = 4 and
# N = 2 bench lines, phase trace, NVLink byte counters around a block of peer-memory / NCCL traversals.
set -u
OUT=gpurun_out/r2k
mkdir -p "$OUT"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl_two_or_more or cxx_nccl" 2>&1 | tail -4 > "$OUT/pytest_multi.txt"
tail -2 "$OUT/pytest_multi.txt"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29579"
TR2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580"
( time $TR4 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline ) > "$OUT/bench_n4.json" 2> "$OUT/bench_n4.err"
tail -1 "$OUT/bench_n4.json" | cut -c1-700
( time $TR2 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
tail -1 "$OUT/bench_n2.json" | cut -c1-500
( B2G_TRACE=1 $TR4 bench.py --gpus 4 --steps 1 --warmup 3 --no-cpu-baseline ) > "$OUT/trace_n4.json" 2> "$OUT/trace_n4.err"
grep "b2g-p2p\] rank 0 phases" "$OUT/trace_n4.err" | tail -1 | cut -c1-700
# NVLink data counters of GPU 0 around 200 traversals with each exchange (KiB; the counters are cumulative)
nvidia-smi nvlink -gt d -i 0 > "$OUT/nvlink_before_p2p.txt" 2>&1
$TR4 bench.py --gpus 4 --steps 200 --warmup 3 --no-cpu-baseline --exchange p2p > "$OUT/nvlink_run_p2p.json" 2> /dev/null
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
    print("GPU 0 NVLink over the whole bench.py run (both exchanges, warm-ups, parity gather included): tx %.1f MB rx %.1f MB" % ((b[0]-a[0])/1024, (b[1]-a[1])/1024))
except Exception as ex:
    print("nvlink counters unavailable:", ex, open(f"{out}/nvlink_after_p2p.txt").read()[:300])
PY
ls -la "$OUT"
