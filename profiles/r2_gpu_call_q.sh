#!/bin/bash
# Round 2, GPU call Q (2 GPUs): bfs / sssp / pr over a multi_context_t of two DISTINCT devices (peer-memory BFS
# exchange; SSSP / PageRank over peer loads + host barriers), three ranks on two devices, and B2G_EXCHANGE=nccl.
set -u
OUT=gpurun_out/r2q
mkdir -p "$OUT"
( time examples/bin/multi_context_selftest 18 0 1 ) > "$OUT/multi_context_2dev.txt" 2>&1
grep -c "wrong depths: 0" "$OUT/multi_context_2dev.txt"; grep -E "sssp|pagerank|ALL OK|FAILED|real" "$OUT/multi_context_2dev.txt" | cut -c1-220
( CUDA_MODULE_LOADING=EAGER B2G_P2P_TIMEOUT_MS=20000 examples/bin/multi_context_selftest 16 0 1 0 ) > "$OUT/multi_context_3ranks_2dev.txt" 2>&1
tail -3 "$OUT/multi_context_3ranks_2dev.txt" | cut -c1-220
( B2G_EXCHANGE=nccl examples/bin/multi_context_selftest 16 0 1 ) > "$OUT/multi_context_2dev_nccl.txt" 2>&1
tail -2 "$OUT/multi_context_2dev_nccl.txt" | cut -c1-220
ls -la "$OUT"
