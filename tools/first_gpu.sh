#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -k "not conv_tc and not fp16" 2>&1 | tail -60 | tee gpurun_out/first_tests.log
