#!/bin/bash
mkdir -p gpurun_out
LIDOG_DBG=0 timeout 200 python tools/trace_units.py 2>&1 | tee gpurun_out/g6_trace_full.txt | tail -75
LIDOG_DBG=103 timeout 200 python tools/trace_units.py 2>&1 | tee gpurun_out/g6_trace_skeleton.txt | tail -30
