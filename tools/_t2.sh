set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -x -q -k "fused_attention" 2>&1 | tail -12
timeout 200 python tools/probs_bench.py
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -5
timeout 200 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/t2_train.json 2> gpurun_out/t2_train.err; tail -2 gpurun_out/t2_train.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t2_train.json').read().strip().splitlines()[-1]); print('train', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_probs -c 1 --launch-skip 9 -o gpurun_out/r02_ds16 -f python tools/probs_bench.py > /dev/null 2>&1
