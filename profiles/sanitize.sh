#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity suite (race detection / sanitizers).
# memcheck: out-of-bounds / misaligned accesses in every kernel the selected tests launch; racecheck: shared-memory hazards in
# the streaming kernels (mbarrier-synchronised bulk-copy rings, tcgen05 kernels, the per-column radix select); synccheck:
# barrier / mbarrier misuse. Usage (GPU box): bash profiles/sanitize.sh [memcheck|racecheck|synccheck|all]
set -u
TOOL=${1:-all}
WIDE="(test_full_run_matches_reference and 2d_full]) or (test_sparse_estep_matches_float64_oracle and 0]) or (test_voxel_data_device_matches_host and 2-float32) or test_gene_cost_kl_matches_oracle or (test_kwargs_surface and large_K)"
NARROW="(test_single_estep_matches_float64_oracle and 2d_full and 0-) or (test_sparse_estep_matches_float64_oracle and 0]) or test_gene_cost_kl_matches_oracle"
run() {  # tool, pytest args...
  local tool=$1; shift
  echo "=== compute-sanitizer --tool $tool : $*"
  compute-sanitizer --tool "$tool" --error-exitcode 9 --launch-timeout 0 python -m pytest -q -x "$@" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Race reported|hazard|Error:|error" | tail -12
  echo "exit code: ${PIPESTATUS[0]}"
}
if [ "$TOOL" = memcheck ] || [ "$TOOL" = all ]; then
  run memcheck tests/test_gpu_parity.py -k "$WIDE"
  run memcheck tests/test_gpu_gram.py -k "64-7000 or 257-4100 or 15-5000 or 3-900"
  run memcheck tests/test_gpu_shard.py -k "1]"
fi
if [ "$TOOL" = racecheck ] || [ "$TOOL" = all ]; then
  run racecheck tests/test_gpu_parity.py -k "$NARROW"
  run racecheck tests/test_gpu_gram.py -k "64-7000"
fi
if [ "$TOOL" = synccheck ] || [ "$TOOL" = all ]; then
  run synccheck tests/test_gpu_parity.py -k "$NARROW"
  run synccheck tests/test_gpu_gram.py -k "64-7000"
fi
