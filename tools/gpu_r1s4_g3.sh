#!/bin/bash
mkdir -p gpurun_out
for d in 103 231; do
  echo "== fwd LIDOG_DBG=$d" | tee -a gpurun_out/g3_diag.txt
  LIDOG_DBG=$d timeout 300 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 10 2>&1 | tee -a gpurun_out/g3_diag.txt | cut -c1-100
done
for d in 231; do
  echo "== roles LIDOG_DBG=$d" | tee -a gpurun_out/g3_roles.txt
  LIDOG_DBG=$d timeout 300 python tools/prof_roles.py 2>&1 | tee -a gpurun_out/g3_roles.txt | head -17
done
