#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/conv_bench.py --cases all 2>&1 | tee gpurun_out/conv_bench.log | tail -80
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 1 -c 1 -o gpurun_out/prof_gemm_g1 \
  python tools/conv_bench.py --cases top --gather 1 --only fwd --reps 1 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 1 -c 1 -o gpurun_out/prof_gemm_g0 \
  python tools/conv_bench.py --cases top --gather 0 --only fwd --reps 1 > gpurun_out/ncu0.log 2>&1; tail -2 gpurun_out/ncu0.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad_tc -s 1 -c 1 -o gpurun_out/prof_wgrad_g1 \
  python tools/conv_bench.py --cases top --gather 1 --only wgrad --reps 1 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ls -la gpurun_out
