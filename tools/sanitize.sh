#!/bin/bash
# compute-sanitizer passes over the tcgen05 / TMA / mbarrier kernels on small shapes.  Run on a GPU box:
#   bash tools/sanitize.sh            -> logs under gpurun_out/sanitizer_*.log, summaries copied to profiles/ by hand
mkdir -p gpurun_out
SEL='tests/test_gpu_kernels.py::test_gemm_concat_projection_residual_layernorm_mask tests/test_gpu_kernels.py::test_gemm_conv3_relu_then_conv3_layernorm tests/test_gpu_kernels.py::test_mha_varlen tests/test_gpu_train.py::test_wgrad_conv_and_concat tests/test_gpu_train.py::test_layernorm_bwd_vectorised_and_fused_column_sums tests/test_gpu_kernels.py::test_stft_mel_golden_and_oracle'
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 --print-limit 20 python -m pytest $SEL -m gpu -q -x -k "not simt" > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -c 'ERROR SUMMARY' gpurun_out/sanitizer_$tool.log) summaries"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_$tool.log | tail -4
done
