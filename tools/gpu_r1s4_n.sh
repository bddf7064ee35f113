#!/bin/bash
# session 4, call N: deeper operand ring (16 id slots, 2 weight stages) vs the validated ring
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/n_tests.log
echo "== fwd ring new" | tee -a gpurun_out/n_sweep.txt
timeout 60 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 6 2>&1 | tee -a gpurun_out/n_sweep.txt | cut -c1-100
echo "== fwd ring old" | tee -a gpurun_out/n_sweep.txt
LIDOG_G2_RING=0 timeout 60 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 6 2>&1 | tee -a gpurun_out/n_sweep.txt | cut -c1-100
timeout 90 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; tail -1 gpurun_out/n_bench.err | cut -c1-120
