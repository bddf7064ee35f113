#!/bin/bash
# session 4, call M: ncu source-level stall view of the production k_gemm2 on the block8 layer shape
mkdir -p gpurun_out /tmp/ncu
timeout 170 ncu --set full --import-source on --clock-control none -k regex:k_gemm2 -s 1 -c 1 -o /tmp/ncu/gemm2src \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only fwd --reps 1 > gpurun_out/m_ncu.log 2>&1
timeout 60 ncu -i /tmp/ncu/gemm2src.ncu-rep --page source --csv > gpurun_out/m_gemm2_source.csv 2> gpurun_out/m_src.err
ls -la gpurun_out/m_gemm2_source.csv; head -c 600 gpurun_out/m_gemm2_source.csv
