#!/bin/bash
# ncu --set full captures: the dominant conv kernels on the block8 shape (ts1, 96->96) and the HBM-bound kernels of one step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm2 -s 1 -c 1 -o gpurun_out/s3_prof_gemm2 \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only fwd --reps 1 > gpurun_out/s3_ncu1.log 2>&1; tail -2 gpurun_out/s3_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad2 -s 1 -c 1 -o gpurun_out/s3_prof_wgrad2 \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only wgrad --reps 1 > gpurun_out/s3_ncu2.log 2>&1; tail -2 gpurun_out/s3_ncu2.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_bev_pool|k_bn_apply|k_bn_bwd_apply|k_bn_stats|k_bn_bwd_stats|k_neighbors|k_conv_c1' -c 40 -o gpurun_out/s3_prof_hbm \
  python bench.py --ncu > gpurun_out/s3_ncu3.log 2>&1; tail -2 gpurun_out/s3_ncu3.log
ls -la gpurun_out/*.ncu-rep
