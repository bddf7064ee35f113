#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/v2_tests.log
timeout 600 python tools/conv_bench.py --cases all --gather 0,2 2>&1 | tee gpurun_out/conv_bench_v2.log | tail -3
timeout 300 python tools/prof_roles.py 2>&1 | tee gpurun_out/roles.log
