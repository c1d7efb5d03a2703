mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "fused_attention" 2>&1 | tail -4
timeout 200 python tools/probs_bench.py
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_probs -c 1 --launch-skip 26 -o gpurun_out/r02_ds16 -f python tools/probs_bench.py > gpurun_out/_ncu.log 2>&1
