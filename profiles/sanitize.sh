#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity suite (race detection / sanitizers).
# memcheck: out-of-bounds / misaligned accesses in every kernel the selected tests launch; racecheck: shared-memory hazards
# in the streaming kernels (mbarrier-synchronised rings are reported as hazards only if a barrier is really missing).
# Usage (GPU box): bash profiles/sanitize.sh [memcheck|racecheck|synccheck] ["pytest -k expression"]
set -u
TOOL=${1:-memcheck}
SEL=${2:-"(test_full_run_matches_reference and 2d_full]) or (test_sparse_estep_matches_float64_oracle and 0]) or (test_voxel_data_device_matches_host and 2-float32) or test_gene_cost_kl_matches_oracle"}
compute-sanitizer --tool "$TOOL" --error-exitcode 9 --launch-timeout 0 \
  python -m pytest tests/test_gpu_parity.py -q -x -k "$SEL" 2>&1 | tail -25
echo "compute-sanitizer $TOOL exit code: ${PIPESTATUS[0]}"
