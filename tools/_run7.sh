set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_train.py tests/test_gpu_aligner.py tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2g_tests.log
tail -6 gpurun_out/r2g_tests.log
python bench.py --mode train --steps 20 --warmup 5 > gpurun_out/r2g_train.json 2> gpurun_out/r2g_train.err
tail -3 gpurun_out/r2g_train.err
python - <<'PY'
import json
for f in ('gpurun_out/r2g_train.json',):
    try:
        t=json.loads(open(f).read().strip().splitlines()[-1]); print(f, 'train', t['value'], t['ms_per_step'], t['e2e']['value'], t['loss'], t['gpu_launches'])
    except Exception as e: print(f, e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2g_launches_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-graphs > /dev/null 2>&1
