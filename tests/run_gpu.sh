#!/bin/bash
# Run on the GPU box (via gpurun): build check, then the gpu-marked tests file by file with timeouts; logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  name=$(basename $f .py)
  timeout 900 python -m pytest $f -m gpu -q -x --timeout 600 2>&1 | tail -60 > gpurun_out/$name.log
  echo "== $name exit ${PIPESTATUS[0]}" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
