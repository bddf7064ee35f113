#!/bin/bash
# session 4, call L: what bounds the LEAN k_gemm2 (DBG instantiation, switches: 16 = no proxy fence [baseline of the DBG build],
# +1 no weight loads, +2 no row gathers, +4 no MMAs)
mkdir -p gpurun_out
for d in 16 17 18 20 23; do
  echo "== fwd LIDOG_DBG=$d" | tee -a gpurun_out/l_diag.txt
  LIDOG_DBG=$d timeout 200 python tools/conv_bench.py --cases net --gather 2 --sorted 1 --only fwd --reps 10 2>&1 | tee -a gpurun_out/l_diag.txt | cut -c1-100
done
