#!/bin/bash
# session 4, call F (1 GPU): tests with failure detail, quick bench, launch list, k_gemm2 diagnostics
# (LIDOG_DBG: 1 = no weight loads, 2 = no row gathers, 4 = no MMAs)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py -m gpu -q -x 2>&1 | tail -40 | tee gpurun_out/f_trainer_test.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_trainer.py 2>&1 | tail -4 | tee gpurun_out/f_gpu_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -1 gpurun_out/f_bench.err | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/f_launches.csv python bench.py --ncu > gpurun_out/f_ncu_launch.log 2>&1; wc -l gpurun_out/f_launches.csv
python tools/launch_summary.py gpurun_out/f_launches.csv > gpurun_out/f_launch_summary.txt 2>&1; head -36 gpurun_out/f_launch_summary.txt
for d in 1 2 4; do
  echo "== fwd LIDOG_DBG=$d" | tee -a gpurun_out/f_diag.txt
  LIDOG_DBG=$d timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 10 2>&1 | tee -a gpurun_out/f_diag.txt | cut -c1-200
done
du -sh gpurun_out
