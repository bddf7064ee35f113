#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/prof_roles.py; echo "=== no proxy fence"; LIDOG_DBG=16 timeout 300 python tools/prof_roles.py
for d in 0 16; do echo "DBG=$d"; LIDOG_DBG=$d timeout 300 python tools/conv_bench.py --cases top --gather 2 --only fwd 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['case'], d['ms'], d['sm_cycles_per_unit'])
"
done) 2>&1 | tee gpurun_out/dbg.log
