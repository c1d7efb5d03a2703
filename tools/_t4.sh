mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12
timeout 200 python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/t4_train.json 2> gpurun_out/t4_train.err; tail -2 gpurun_out/t4_train.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/t4_train.json').read().strip().splitlines()[-1]); print('train', d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'))
PY
