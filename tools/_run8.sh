set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_audio_inverse.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2h_tests.log
tail -30 gpurun_out/r2h_tests.log
