#!/bin/bash
# Round 2, GPU call Q (4 GPUs): the final code on distinct devices -- bfs / sssp / pr over a 4-device multi_context_t
# (peer-memory BFS exchange, thread exchange for SSSP / PageRank, then B2G_EXCHANGE=nccl for BFS), the torchrun
# worker at 4 ranks, and the N = 4 bench line.
set -u
OUT=gpurun_out/r2q
mkdir -p "$OUT"
( time examples/bin/multi_context_selftest 18 0 1 2 3 ) > "$OUT/multi_context_4dev.txt" 2>&1
tail -12 "$OUT/multi_context_4dev.txt" | cut -c1-200
( B2G_EXCHANGE=nccl examples/bin/multi_context_selftest 16 0 1 2 3 ) > "$OUT/multi_context_4dev_nccl.txt" 2>&1
tail -2 "$OUT/multi_context_4dev_nccl.txt" | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl_two_or_more" 2>&1 | tail -6 > "$OUT/pytest_multi.txt"
tail -3 "$OUT/pytest_multi.txt"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29579"
( time $TR4 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline ) > "$OUT/bench_n4.json" 2> "$OUT/bench_n4.err"
tail -1 "$OUT/bench_n4.json" | cut -c1-600
ls -la "$OUT"
