#!/bin/bash
# compute-sanitizer over a slice of the GPU parity tests (memcheck + racecheck + synccheck)
SEL="kitten_4096 or tiny or pyramid_sizes or tma_staged or overflow or taskcull or hostile or big_meshes or decode or without_prepared"
for tool in ${1:-memcheck racecheck synccheck}; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error:|hazard" | head -8
done
