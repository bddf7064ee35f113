#!/bin/bash
# Round 2, first GPU call: re-validate the round-1 state, then check the paths that were written after the round-1
# GPU budget was spent:
#   1. full GPU parity suite + smoke + bench (must match profiles/r01_v6_bench.json within box variance)
#   2. the experimental two-issuer k_gemm2 (LIDOG_G2_MMA2=1) after the parity-aliasing fix: conv parity tests, then
#      the layer-shape sweep against the single-issuer default
#   3. the ring rule of the last commit on the whole training step (only the conv tests saw it on hardware)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2a_gpu_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -1 gpurun_out/r2a_bench.err | cut -c1-200
LIDOG_G2_MMA2=1 timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2a_tests_mma2.log
for cfg in "LIDOG_G2_MMA2=0" "LIDOG_G2_MMA2=1"; do
  echo "== fwd,dgrad $cfg" | tee -a gpurun_out/r2a_sweep_mma2.txt
  env $cfg timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd,dgrad --reps 10 2>&1 | tee -a gpurun_out/r2a_sweep_mma2.txt | cut -c1-120
done
LIDOG_G2_MMA2=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_mma2.json 2> gpurun_out/r2a_bench_mma2.err; tail -1 gpurun_out/r2a_bench_mma2.err | cut -c1-200
# 4. BASELINE configs[3]: micro-benchmark sweep over the point count on random-plane clouds (fwd / dgrad / wgrad)
for n in 10000 100000 1000000 2000000; do
  echo "== points $n" | tee -a gpurun_out/r2a_microbench.txt
  timeout 300 python tools/conv_bench.py --points $n --cases all --gather 2 --sorted 1 --reps 10 2>&1 | tee -a gpurun_out/r2a_microbench.txt | cut -c1-120
done
