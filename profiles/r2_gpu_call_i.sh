#!/bin/bash
# Round 2, GPU call I (8 GPUs): the bench exactly as the driver launches it at N = 8 and N = 4 (strong scaling of ONE
# RMAT-26 traversal, peer-memory exchange, the C++ NCCL loop timed beside it, gathered-depth parity), bfs::run over an
# 8-device multi_context_t, per-phase trace.
set -u
OUT=gpurun_out/r2i
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt"
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29578"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29579"
( time $TR8 bench.py --gpus 8 --steps 20 --warmup 5 ) > "$OUT/bench_n8.json" 2> "$OUT/bench_n8.err"
tail -3 "$OUT/bench_n8.err"; tail -1 "$OUT/bench_n8.json" | cut -c1-1300
( time $TR4 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline ) > "$OUT/bench_n4.json" 2> "$OUT/bench_n4.err"
tail -1 "$OUT/bench_n4.json" | cut -c1-700
( B2G_TRACE=1 $TR8 bench.py --gpus 8 --steps 1 --warmup 3 --no-cpu-baseline ) > "$OUT/trace_n8.json" 2> "$OUT/trace_n8.err"
grep "b2g-p2p\] rank 0 phases" "$OUT/trace_n8.err" | tail -1 | cut -c1-700
grep "b2g-nccl\] rank 0" "$OUT/trace_n8.err" | tail -7
( timeout 300 examples/bin/multi_context_selftest 20 0 1 2 3 4 5 6 7 ) > "$OUT/multi_context_8.txt" 2>&1; tail -3 "$OUT/multi_context_8.txt"
ls -la "$OUT"
