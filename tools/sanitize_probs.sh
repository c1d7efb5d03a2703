#!/bin/bash
# compute-sanitizer over the attention-probability kernels (attn_probs_tc.cu, both modes) on the smallest test shape.
mkdir -p gpurun_out
SEL='tests/test_gpu_train.py'
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-80} compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 10 python -m pytest $SEL -m gpu -q -x -k "fused_attention and 130" > gpurun_out/sanitizer_probs_$tool.log 2>&1
  echo "== $tool: rc $?"; grep -E "ERROR SUMMARY|passed|failed|Race|hazard" gpurun_out/sanitizer_probs_$tool.log | sort | uniq -c | tail -6
done
