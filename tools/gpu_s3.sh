#!/bin/bash
# session-3 run D: acc sets A/B, channels_last A/B, launch list
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tee gpurun_out/s3d_tests.log
for sets in 1 2; do
  echo "== LIDOG_ACC_SETS=$sets"
  LIDOG_ACC_SETS=$sets timeout 600 python tools/conv_bench.py --cases all --gather 2 --sorted 1 --only fwd,dgrad 2>&1 | tee gpurun_out/s3d_conv_bench_sets$sets.log | tail -2
done
timeout 300 python tools/prof_roles.py 2>&1 | tee gpurun_out/s3d_roles.log | grep -A16 "ts1 96"
for cl in 0 1; do
  LIDOG_BEV_CHANNELS_LAST=$cl timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/s3d_bench_cl$cl.err > gpurun_out/s3d_bench_cl$cl.json
  tail -1 gpurun_out/s3d_bench_cl$cl.err | cut -c1-120
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s3d_launches.csv python bench.py --ncu > gpurun_out/s3d_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s3d_launches.csv > gpurun_out/s3d_launch_summary.txt 2>&1; head -50 gpurun_out/s3d_launch_summary.txt
