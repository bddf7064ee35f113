#!/bin/bash
# session-3 run H: fused BN + BEV spans
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) 2>&1 | tee gpurun_out/s3i_tests.log
for f in 0 1; do
LIDOG_FUSED_BN=$f timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2> gpurun_out/s3i_bench_f$f.err > gpurun_out/s3i_bench_f$f.json
tail -1 gpurun_out/s3i_bench_f$f.err | cut -c1-160
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/s3i_launches.csv python bench.py --ncu > gpurun_out/s3i_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/s3i_launches.csv > gpurun_out/s3i_launch_summary.txt 2>&1; head -40 gpurun_out/s3i_launch_summary.txt
