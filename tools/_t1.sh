set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "fused_attention" 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_dp.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -25
timeout 200 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/t1_train.json 2> gpurun_out/t1_train.err; tail -2 gpurun_out/t1_train.err
TTSB_NO_FUSED_PROBS=1 timeout 200 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/t1_train_nofuse.json 2>/dev/null
python - <<'PY'
import json
for f in ('t1_train','t1_train_nofuse'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
    except Exception as e: print(f, e)
PY
