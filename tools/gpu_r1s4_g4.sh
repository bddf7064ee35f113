#!/bin/bash
mkdir -p gpurun_out
for d in 103 359 615 871; do
  echo "== roles LIDOG_DBG=$d" | tee -a gpurun_out/g4_roles.txt
  LIDOG_DBG=$d timeout 300 python tools/prof_roles.py 2>&1 | tee -a gpurun_out/g4_roles.txt | head -18
done
