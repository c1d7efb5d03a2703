#!/bin/bash
# Run every test id of the given file in its own process (a sticky CUDA error must not cascade); logs in gpurun_out/.
mkdir -p gpurun_out
f=$1
name=$(basename $f .py)
ids=$(python -m pytest $f -m gpu --collect-only -q 2>/dev/null | grep "::")
: > gpurun_out/${name}_each.log
for id in $ids; do
  echo "=== $id" >> gpurun_out/${name}_each.log
  CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest "$id" -m gpu -q -x --timeout 250 2>&1 | grep -E "passed|failed|Error|error|assert|worst|rel" | head -12 >> gpurun_out/${name}_each.log
done
cat gpurun_out/${name}_each.log
