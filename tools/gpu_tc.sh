#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/tc_probe.py 2>&1 | tee gpurun_out/tc_probe.log
