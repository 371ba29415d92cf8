#!/bin/bash
# compute-sanitizer over a small slice of the GPU parity tests (memcheck + racecheck + synccheck)
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "kitten_4096 or tiny or pyramid_sizes or tma_staged or overflow or taskcull" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error:" | head -8
done
