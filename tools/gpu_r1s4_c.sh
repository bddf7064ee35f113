#!/bin/bash
# session 4, call C: full validation of the round state + profiles (tests, smoke, bench both arms, launch list, ncu full)
mkdir -p gpurun_out
nproc > gpurun_out/c_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/c_host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/c_host.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/c_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/c_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; tail -2 gpurun_out/c_bench.err | cut -c1-200; cut -c1-300 gpurun_out/c_bench.json
timeout 300 python tools/layer_table.py > gpurun_out/c_layer_table.txt 2> gpurun_out/c_layer_table.err; head -3 gpurun_out/c_layer_table.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/c_launches.csv python bench.py --ncu > gpurun_out/c_ncu_launch.log 2>&1; tail -1 gpurun_out/c_ncu_launch.log; wc -l gpurun_out/c_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gemm2 -s 1 -c 1 -o gpurun_out/c_prof_gemm2 \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only fwd --reps 1 > gpurun_out/c_ncu1.log 2>&1; tail -1 gpurun_out/c_ncu1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_wgrad2 -s 1 -c 1 -o gpurun_out/c_prof_wgrad2 \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only wgrad --reps 1 > gpurun_out/c_ncu2.log 2>&1; tail -1 gpurun_out/c_ncu2.log
timeout 900 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_bev_pool|k_bn_apply|k_bn_bwd_apply|k_bn_stats|k_bn_bwd_stats|k_neighbors|k_conv_c1|k_head' -c 48 -o gpurun_out/c_prof_hbm \
  python bench.py --ncu > gpurun_out/c_ncu3.log 2>&1; tail -1 gpurun_out/c_ncu3.log
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err; tail -2 gpurun_out/c_bench_ref.err; cut -c1-300 gpurun_out/c_bench_ref.json
ls -la gpurun_out/*.ncu-rep
