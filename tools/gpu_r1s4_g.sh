#!/bin/bash
# session 4, call G (1 GPU): trainer test, k_gemm2 skeleton diagnostics
# LIDOG_DBG bits: 1 no weight loads, 2 no row gathers, 4 no MMAs, 32 no result stores, 64 no row-id copies
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py -m gpu -q -x 2>&1 | tail -30 | tee gpurun_out/g_trainer_test.log
for d in 0 7 32 64 39 71 103; do
  echo "== fwd LIDOG_DBG=$d" | tee -a gpurun_out/g_diag.txt
  LIDOG_DBG=$d timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 10 2>&1 | tee -a gpurun_out/g_diag.txt | cut -c1-100
done
