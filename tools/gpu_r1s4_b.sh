#!/bin/bash
# session 4, call B: set-bit offset walk, L2 prefetch by the id warp (OPT bit 4), wgrad chunk rule
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/b_tests.log
for cfg in "LIDOG_G2_OPT=7" "LIDOG_G2_OPT=3" "LIDOG_G2_OPT=7 LIDOG_ACC_SETS=1" "LIDOG_G2_OPT=7 LIDOG_G2_SB=2"; do
  echo "== fwd $cfg" | tee -a gpurun_out/b_sweep.txt
  env $cfg timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd,dgrad --reps 10 2>&1 | tee -a gpurun_out/b_sweep.txt | cut -c1-230
done
echo "== wgrad default-rule" | tee -a gpurun_out/b_sweep.txt
timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only wgrad --reps 10 2>&1 | tee -a gpurun_out/b_sweep.txt | cut -c1-230
timeout 300 python tools/prof_roles.py > gpurun_out/b_roles.txt 2>&1; tail -64 gpurun_out/b_roles.txt
timeout 300 python tools/layer_table.py > gpurun_out/b_layer_table.txt 2> gpurun_out/b_layer_table.err; tail -3 gpurun_out/b_layer_table.err; head -12 gpurun_out/b_layer_table.txt | cut -c1-220
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; tail -2 gpurun_out/b_bench.err | cut -c1-300; cut -c1-400 gpurun_out/b_bench.json
