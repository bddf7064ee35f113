#!/bin/bash
# session 4, call A: correctness of the changed kernels, per-layer table, pipeline-shape sweeps, quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/a_tests.log
timeout 300 python tools/layer_table.py > gpurun_out/a_layer_table.txt 2> gpurun_out/a_layer_table.err; tail -3 gpurun_out/a_layer_table.err; head -40 gpurun_out/a_layer_table.txt
for cfg in "LIDOG_G2_OPT=3" "LIDOG_G2_OPT=0" "LIDOG_G2_OPT=1" "LIDOG_G2_OPT=2" "LIDOG_G2_OPT=3 LIDOG_G2_PC=1 LIDOG_G2_SB=4" "LIDOG_G2_OPT=3 LIDOG_G2_PC=2 LIDOG_G2_SB=4" "LIDOG_G2_OPT=3 LIDOG_DBG=16"; do
  echo "== fwd $cfg" | tee -a gpurun_out/a_sweep.txt
  env $cfg timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 10 2>&1 | tee -a gpurun_out/a_sweep.txt | cut -c1-230
done
for cfg in "LIDOG_WG_CTAS=592" "LIDOG_WG_CTAS=296" "LIDOG_WG_CTAS=148"; do
  echo "== wgrad $cfg" | tee -a gpurun_out/a_sweep.txt
  env $cfg timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only wgrad --reps 10 2>&1 | tee -a gpurun_out/a_sweep.txt | cut -c1-230
done
timeout 300 python tools/prof_roles.py > gpurun_out/a_roles.txt 2>&1; tail -50 gpurun_out/a_roles.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; tail -3 gpurun_out/a_bench.err; cat gpurun_out/a_bench.json
