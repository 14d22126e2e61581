#!/bin/bash
# Round 2, GPU call T (2 GPUs): the final build on two devices -- multi-GPU test file (torchrun worker, 63 checks),
# bfs / sssp / pr over a two-device multi_context_t, the N = 2 bench line exactly as the driver launches it.
set -u
OUT=gpurun_out/r2t
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_examples.py -m gpu -q 2>&1 | tail -8 > "$OUT/pytest_multi.txt"
tail -3 "$OUT/pytest_multi.txt"
( examples/bin/multi_context_selftest 17 0 1 ) > "$OUT/multi_context_2dev.txt" 2>&1; tail -2 "$OUT/multi_context_2dev.txt" | cut -c1-200
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29580 bench.py --gpus 2 --steps 20 --warmup 5 ) > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
tail -4 "$OUT/bench_n2.err"; tail -1 "$OUT/bench_n2.json" | cut -c1-500
ls -la "$OUT"
