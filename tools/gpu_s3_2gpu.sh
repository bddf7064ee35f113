#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/s3j_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/syncbn_check.py > gpurun_out/s3_syncbn.log 2>&1
grep -v "^\[rank\|^W1\|^\*\*\*" gpurun_out/s3_syncbn.log | tail -6
for peer in 1 0; do
LIDOG_PEER_SYNCBN=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$peer bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/s3j_bench_2gpu_peer$peer.err > gpurun_out/s3j_bench_2gpu_peer$peer.json
grep "bench\]\|Warn\|warn" gpurun_out/s3j_bench_2gpu_peer$peer.err | cut -c1-200
done
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/s3j_bench_1gpu.err > gpurun_out/s3j_bench_1gpu.json; grep "bench\]" gpurun_out/s3j_bench_1gpu.err | cut -c1-150
