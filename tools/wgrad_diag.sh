#!/bin/bash
# Where does k_wgrad2's time go?  Layer-shape sweep (wgrad only) with parts of the kernel switched off (LIDOG_WG_DBG:
# 1 = no dY gathers, 2 = no X gathers, 4 = no MMAs, 8 = no partial stores) and with ring / CTA-count variants.
for cfg in "LIDOG_WG_DBG=0" "LIDOG_WG_DBG=1" "LIDOG_WG_DBG=2" "LIDOG_WG_DBG=4" "LIDOG_WG_DBG=7" "LIDOG_WG_DBG=15" \
           "LIDOG_WG_SB=2" "LIDOG_WG_CTAS=148"; do
  echo "== $cfg"
  env $cfg timeout 200 python tools/conv_bench.py --cases net --sorted 1 --only wgrad --reps 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('  %-28s %7.4f ms %6.1f TF/s' % (d['case'], d['ms'], d['alg_tflops']))
"
done
