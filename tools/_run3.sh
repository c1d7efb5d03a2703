set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2c_tests.log
tail -8 gpurun_out/r2c_tests.log
python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/r2c_train.json 2> gpurun_out/r2c_train.err
python bench.py --mode train --steps 20 --warmup 5 --no-graphs > gpurun_out/r2c_train_eager.json 2> gpurun_out/r2c_train_eager.err
tail -3 gpurun_out/r2c_train.err
python - <<'PY'
import json
for f in ('gpurun_out/r2c_train.json','gpurun_out/r2c_train_eager.json'):
    try:
        t=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'train', t['value'], t['ms_per_step'], t['e2e']['value'], t['loss'], t['gpu_launches'])
    except Exception as e: print(f, e)
PY
