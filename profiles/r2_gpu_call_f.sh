#!/bin/bash
# Round 2, GPU call F (2 GPUs): the multi-GPU paths on real devices -- torchrun worker (NCCL via torch, the C++ NCCL
# level loop, the peer-memory exchange; 58 checks), bfs::run over a two-device multi_context_t (peer memory and
# B2G_EXCHANGE=nccl), the N = 2 bench line (strong scaling, both exchanges, gathered-depth parity), phase traces.
set -u
OUT=gpurun_out/r2f
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt"
# single-GPU sanity of the newest pull kernel (K1 with 4 vertices per lane in flight) before the multi-GPU part
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernel_switches.py -m gpu -q -x 2>&1 | tail -4 > "$OUT/pytest_single.txt"
tail -2 "$OUT/pytest_single.txt"
python bench.py --workload bfs_do_rmat26 --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > "$OUT/do26.json"
python -c "
import json; j=json.load(open('$OUT/do26.json')); c=j['config']
print('do26 %.3f ms/step' % j['ms_per_step'], c['level_kernel_ms'], 'outside %.3f' % c['outside_kernels_frac'])"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "nccl_two_or_more or cxx_nccl" 2>&1 | tail -15 > "$OUT/pytest_multi.txt"
tail -4 "$OUT/pytest_multi.txt"
( timeout 300 examples/bin/multi_context_selftest 19 0 1 ) > "$OUT/multi_context_p2p.txt" 2>&1; tail -3 "$OUT/multi_context_p2p.txt"
( B2G_EXCHANGE=nccl timeout 300 examples/bin/multi_context_selftest 19 0 1 ) > "$OUT/multi_context_nccl.txt" 2>&1; tail -3 "$OUT/multi_context_nccl.txt"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577"
( time $TR bench.py --gpus 2 --steps 10 --warmup 3 ) > "$OUT/bench_n2_p2p.json" 2> "$OUT/bench_n2_p2p.err"
tail -3 "$OUT/bench_n2_p2p.err"; tail -1 "$OUT/bench_n2_p2p.json" | cut -c1-1500
( time $TR bench.py --gpus 2 --steps 10 --warmup 3 --exchange nccl --no-cpu-baseline ) > "$OUT/bench_n2_nccl.json" 2> "$OUT/bench_n2_nccl.err"
tail -3 "$OUT/bench_n2_nccl.err"; tail -1 "$OUT/bench_n2_nccl.json" | cut -c1-900
( B2G_TRACE=1 $TR bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline ) > "$OUT/trace_n2.json" 2> "$OUT/trace_n2.err"
grep "b2g-p2p\] rank 0 phases\|b2g-nccl\] rank 0" "$OUT/trace_n2.err" | tail -12
ls -la "$OUT"
