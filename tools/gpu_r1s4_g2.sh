#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py -m gpu -q 2>&1 | grep -E "^E  |passed|failed" | cut -c1-1500 | tee gpurun_out/g2_trainer_test.log
for d in 103 7 0; do
  echo "== roles LIDOG_DBG=$d" | tee -a gpurun_out/g2_roles.txt
  LIDOG_DBG=$d timeout 300 python tools/prof_roles.py 2>&1 | tee -a gpurun_out/g2_roles.txt | head -34
done
