#!/bin/bash
# full GPU validation: parity tests, smoke, bench (ours + reference arm), ncu launch list
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/host.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python bench.py --ncu > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
