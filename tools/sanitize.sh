#!/bin/bash
# compute-sanitizer target: memcheck over whole substeps of all four solver classes (flip, viscous nbflip,
# smoke, fire), the packed and the streamed particle transfers. Run on a GPU box: `gpurun -- bash tools/sanitize.sh`.
# Single-handle tests only: the slab kernels of several ranks wait for each other with a spin limit that the sanitizer's
# slow-down would trip.
set -o pipefail
compute-sanitizer --tool "${1:-memcheck}" --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_substep_gpu.py tests/test_host_solver_gpu.py -x -q -m gpu \
    -k "deterministic or streamed_round_trip or granular or streamed_substeps" 2>&1 | tee gpurun_out/sanitizer.log
