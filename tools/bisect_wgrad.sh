T="tests/test_gpu_fullscale.py::test_fullscale_convolution_matches_float64_oracle"
for cfg in "X=0" "LIDOG_WG_NA4=1" "LIDOG_WG_SB=2" "LIDOG_WG_SA=5" "LIDOG_WG_SA=5 LIDOG_WG_SB=2"; do
  echo "== $cfg"; env $cfg timeout 300 python -m pytest "$T" -m gpu -q -k "2-32-32 or 1-96-96" 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-150
done
