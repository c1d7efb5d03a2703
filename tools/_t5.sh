mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s > gpurun_out/t5_dp.log 2>&1
grep -n "DP vs\|assert \|passed\|failed" gpurun_out/t5_dp.log | cut -c1-700
