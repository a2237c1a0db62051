#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the small GPU tests that exercise the racy-by-design protocols:
# the cooperative grid kernel (phase flags, sparse lists), the 16..31-byte two-half slots, first-wave owner/poll contention.
# Usage (on the GPU box):  tools/sanitize_gpu.sh <tag>     -> gpurun_out/sanitizer_<tag>_{memcheck,racecheck}.txt
tag=${1:-run}
mkdir -p gpurun_out
TESTS='tests/test_gpu_parity.py::test_grid_kernel_random_proper_tables[0] tests/test_gpu_parity.py::test_grid_kernel_random_proper_tables[3] tests/test_gpu_parity.py::test_grid_kernel_sparse_phase_falls_back_for_a_long_late_run tests/test_gpu_parity.py::test_tiles_words_of_16_to_31_bytes_share_prefixes tests/test_gpu_parity.py::test_tiles_dense_isolated_bytes_and_first_wave_contention tests/test_gpu_parity.py::test_truncation_and_padding[0] tests/test_gpu_parity.py::test_bpe_random[1] tests/test_gpu_parity.py::test_wordpiece_random[1]'
for tool in memcheck racecheck; do
  out=gpurun_out/sanitizer_${tag}_${tool}.txt
  echo "## compute-sanitizer --tool $tool  ($(date -u +%FT%TZ))" > $out
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest -x -q $TESTS -p no:cacheprovider >> $out 2>&1
  echo "## exit code $?" >> $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out | tail -5
done
