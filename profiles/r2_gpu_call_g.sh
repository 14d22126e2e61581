#!/bin/bash
# Round 2, GPU call G (2 GPUs): after the timer fix and the out-degree sum at the pull -> push switch of the two
# multi-GPU loops: multi_context selftests (peer memory, NCCL), torchrun worker, N = 2 bench with both exchanges.
set -u
OUT=gpurun_out/r2g
mkdir -p "$OUT"
( timeout 300 examples/bin/multi_context_selftest 19 0 1 ) > "$OUT/multi_context_p2p.txt" 2>&1; tail -2 "$OUT/multi_context_p2p.txt"
( B2G_EXCHANGE=nccl timeout 300 examples/bin/multi_context_selftest 19 0 1 ) > "$OUT/multi_context_nccl.txt" 2>&1; tail -2 "$OUT/multi_context_nccl.txt"
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_examples.py -m gpu -q -k "nccl_two_or_more or cxx_nccl or multi_device or p2p or caches" 2>&1 | tail -6 > "$OUT/pytest_multi.txt"
tail -3 "$OUT/pytest_multi.txt"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577"
( time $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline ) > "$OUT/bench_n2_p2p.json" 2> "$OUT/bench_n2_p2p.err"
tail -1 "$OUT/bench_n2_p2p.json" | cut -c1-700
( B2G_TRACE=1 $TR bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline ) > "$OUT/trace_n2.json" 2> "$OUT/trace_n2.err"
grep "b2g-p2p\] rank 0 phases" "$OUT/trace_n2.err" | tail -1 | cut -c1-700
grep "b2g-nccl\] rank 0" "$OUT/trace_n2.err" | tail -7
ls -la "$OUT"
