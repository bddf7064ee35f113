#!/bin/bash
# full GPU validation: parity tests, smoke, bench (ours + reference arm)
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/host.txt; free -g | head -2 >> gpurun_out/host.txt
python -m pytest tests -m gpu -q 2>&1 | tail -30 | tee gpurun_out/gpu_tests.log
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
python bench.py --steps 6 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
