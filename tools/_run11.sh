mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k stft 2>&1 | tail -3
TTSB_STFT_V2=1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k stft 2>&1 | tail -3
for i in 1 2; do
python bench.py --mode stft --steps 30 > gpurun_out/r2k_stft_v3.json 2> gpurun_out/r2k_stft_v3.err
TTSB_STFT_V2=1 python bench.py --mode stft --steps 30 > gpurun_out/r2k_stft_v2.json 2> gpurun_out/r2k_stft_v2.err
python - <<'PY'
import json
for f in ('gpurun_out/r2k_stft_v3.json','gpurun_out/r2k_stft_v2.json'):
    try:
        t=json.loads(open(f).read().strip().splitlines()[-1]); print(f, t['ms_per_step'], t['roofline']['frac'], t['cpu_baseline']['max_abs_err_gpu_vs_cpu'])
    except Exception as e: print(f, e, open(f.replace('.json','.err')).read()[-800:])
PY
done
