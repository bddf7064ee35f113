#!/bin/bash
# compute-sanitizer passes over the hand-written kernels: memcheck on the whole small-input parity
# suite of the convolution / BN / map kernels, racecheck (shared-memory hazards) and synccheck on one convolution case
# of every pipeline shape.  The tcgen05 / TMA / mbarrier protocol of k_gemm2 trapped once in round 1; this is the tool
# that looks at it.   usage: bash tools/sanitize.sh gpurun_out/<tag>
O=${1:-gpurun_out/sanitize}
SAN=/usr/local/cuda/bin/compute-sanitizer
K="tests/test_gpu_conv.py::test_conv_tc_ragged_sizes tests/test_gpu_bn.py::test_epilogue_statistics_feed_the_batch_norm"
timeout 900 $SAN --tool memcheck --error-exitcode 3 python -m pytest $K tests/test_gpu_voxel.py -m gpu -q -x > ${O}_memcheck.log 2>&1
echo "memcheck exit $?" | tee -a ${O}_sanitize_summary.txt; grep -E "ERROR SUMMARY|passed|failed" ${O}_memcheck.log | tail -3 | tee -a ${O}_sanitize_summary.txt
timeout 900 $SAN --tool racecheck --error-exitcode 3 python -m pytest "tests/test_gpu_conv.py::test_conv_tc_ragged_sizes" -m gpu -q -x > ${O}_racecheck.log 2>&1
echo "racecheck exit $?" | tee -a ${O}_sanitize_summary.txt; grep -E "RACECHECK SUMMARY|passed|failed" ${O}_racecheck.log | tail -3 | tee -a ${O}_sanitize_summary.txt
timeout 900 $SAN --tool synccheck --error-exitcode 3 python -m pytest "tests/test_gpu_conv.py::test_conv_tc_ragged_sizes" -m gpu -q -x > ${O}_synccheck.log 2>&1
echo "synccheck exit $?" | tee -a ${O}_sanitize_summary.txt; grep -E "ERROR SUMMARY|passed|failed" ${O}_synccheck.log | tail -3 | tee -a ${O}_sanitize_summary.txt
