#!/bin/bash
# session 4, call E (4 GPUs): SyncBN exchange check at 4 ranks, DDP bench at 4 and 2 ranks, 1-rank bench on the same box
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | tee gpurun_out/e_gpus.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 tools/syncbn_check.py > gpurun_out/e_syncbn4.log 2>&1
grep -v "^\[rank\|^W1\|^\*\*\*" gpurun_out/e_syncbn4.log | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 6 --warmup 3 2> gpurun_out/e_bench_4gpu.err > gpurun_out/e_bench_4gpu.json
grep "bench\]\|Warn\|warn\|rror" gpurun_out/e_bench_4gpu.err | cut -c1-200 | head -5; cut -c1-160 gpurun_out/e_bench_4gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/e_bench_2gpu.err > gpurun_out/e_bench_2gpu.json
grep "bench\]\|Warn\|warn\|rror" gpurun_out/e_bench_2gpu.err | cut -c1-200 | head -5; cut -c1-160 gpurun_out/e_bench_2gpu.json
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/e_bench_1gpu.err > gpurun_out/e_bench_1gpu.json; cut -c1-160 gpurun_out/e_bench_1gpu.json
du -sh gpurun_out
