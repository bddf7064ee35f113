#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm2 -s 1 -c 1 -o gpurun_out/prof_gemm2 \
  python tools/conv_bench.py --cases top --gather 2 --only fwd --reps 1 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad2 -s 1 -c 1 -o gpurun_out/prof_wgrad2 \
  python tools/conv_bench.py --cases top --gather 2 --only wgrad --reps 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
