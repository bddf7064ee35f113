#!/bin/bash
# session 4, call K (final validation of the round state): validation + profiles; ncu reports are summarised ON THE BOX (text only travels back: gpurun_out <= 64 MiB)
mkdir -p gpurun_out /tmp/ncu
nproc > gpurun_out/k_host.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/k_host.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/k_host.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/k_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/k_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err; tail -1 gpurun_out/k_bench.err | cut -c1-200
timeout 300 python tools/layer_table.py > gpurun_out/k_layer_table.txt 2> gpurun_out/k_layer_table.err; head -3 gpurun_out/k_layer_table.txt | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/k_launches.csv python bench.py --ncu > gpurun_out/k_ncu_launch.log 2>&1; wc -l gpurun_out/k_launches.csv
python tools/launch_summary.py gpurun_out/k_launches.csv > gpurun_out/k_launch_summary.txt 2>&1; head -30 gpurun_out/k_launch_summary.txt
timeout 600 ncu --set full --clock-control none -k regex:k_gemm2 -s 1 -c 1 -o /tmp/ncu/gemm2 \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only fwd --reps 1 > gpurun_out/k_ncu1.log 2>&1
python tools/ncu_summary.py /tmp/ncu/gemm2.ncu-rep > gpurun_out/k_ncu_gemm2.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_wgrad2 -s 1 -c 1 -o /tmp/ncu/wgrad2 \
  python tools/conv_bench.py --cases top --gather 2 --sorted 1 --only wgrad --reps 1 > gpurun_out/k_ncu2.log 2>&1
python tools/ncu_summary.py /tmp/ncu/wgrad2.ncu-rep > gpurun_out/k_ncu_wgrad2.txt 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k 'regex:k_bev_pool|k_bn_apply|k_bn_bwk_apply|k_bn_stats|k_bn_bwk_stats|k_neighbors|k_conv_c1|k_head' -c 60 -o /tmp/ncu/hbm \
  python bench.py --ncu > gpurun_out/k_ncu3.log 2>&1
python tools/ncu_summary.py /tmp/ncu/hbm.ncu-rep > gpurun_out/k_ncu_hbm_kernels.txt 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/k_bench_ref.json 2> gpurun_out/k_bench_ref.err; cut -c1-200 gpurun_out/k_bench_ref.json
timeout 600 python bench.py --shape nuscenes --batch 16 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_nuscenes.json 2> gpurun_out/k_bench_nuscenes.err; cut -c1-200 gpurun_out/k_bench_nuscenes.json
timeout 600 python bench.py --shape mix3d --batch 8 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_mix3d.json 2> gpurun_out/k_bench_mix3d.err; cut -c1-200 gpurun_out/k_bench_mix3d.json
du -sh gpurun_out
