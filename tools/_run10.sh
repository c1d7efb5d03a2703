mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -q -k stft 2>&1 | tail -5
python -m pytest tests/test_gpu_audio_inverse.py tests/test_gpu_train.py -m gpu -q -x 2>&1 | tail -5
python bench.py --mode stft --steps 20 > gpurun_out/r2j_stft_v2.json 2> gpurun_out/r2j_stft_v2.err
TTSB_STFT_V1=1 python bench.py --mode stft --steps 20 > gpurun_out/r2j_stft_v1.json 2> gpurun_out/r2j_stft_v1.err
python - <<'PY'
import json
for f in ('gpurun_out/r2j_stft_v2.json','gpurun_out/r2j_stft_v1.json'):
    try:
        t=json.loads(open(f).read().strip().splitlines()[-1]); print(f, t['ms_per_step'], t['roofline']['frac'], t['cpu_baseline'])
    except Exception as e: print(f, e, open(f.replace('.json','.err')).read()[-500:])
PY
