#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 2> gpurun_out/i_bench_2gpu.err > gpurun_out/i_bench_2gpu.json
grep "bench\]\|Warn\|warn\|rror" gpurun_out/i_bench_2gpu.err | cut -c1-160 | head -5; python -c "
import json;r=json.loads(open('gpurun_out/i_bench_2gpu.json').read().strip().splitlines()[-1]);print(r['value'],r['ms_per_step'],r['host_issue_ms_per_step'],r['e2e'])"
