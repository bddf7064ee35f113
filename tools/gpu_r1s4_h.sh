#!/bin/bash
# session 4, call H: lean MMA warp (templated production kernel): correctness + sweep + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py tests/test_gpu_trainer.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/h3_tests.log
echo "== fwd,dgrad lean" | tee -a gpurun_out/h3_sweep.txt
timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd,dgrad --reps 10 2>&1 | tee -a gpurun_out/h3_sweep.txt | cut -c1-200
echo "== fwd,dgrad single issuer" | tee -a gpurun_out/h3_sweep.txt; LIDOG_G2_MMA2=0 timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 10 2>&1 | tee -a gpurun_out/h3_sweep.txt | cut -c1-120
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/h3_bench.json 2> gpurun_out/h3_bench.err; tail -1 gpurun_out/h3_bench.err | cut -c1-200
